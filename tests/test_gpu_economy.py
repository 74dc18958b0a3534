"""C4-B households + consumer-goods firms economy (advanced_economic_model.py): CUDA path vs the
NumPy oracle on the same seeds.  float32 trajectories within 1e-5 relative while the model is
finite; the reference model itself degenerates into NaNs within a handful of steps (0/0 once a
firm's inventory covers its demand share, :433-457) and compute_metrics' nan_to_num defaults
become the output -- both sides must agree on that too."""
import numpy as np
import pytest

import jaxabm_b200 as jx
from jaxabm_b200.rules import economy
from oracle import economy as oeco, runtime as ort

pytestmark = pytest.mark.gpu

RTOL = 1e-5     # float32 trajectories (north_star tolerance)
# differences of nearly equal float32 numbers: gdp_growth = (gdp/prev_gdp - 1)*100, inflation likewise.
# One ulp of the ratio (6e-8) is 6e-6 percentage points before any accumulated rounding; the
# tolerance for these is therefore absolute, in percentage points.
ATOL_DIFF = {"gdp_growth": 5e-4, "inflation": 5e-4, "economic_health": 1e-3}


def _cmp_states(dev_states, ora_states, rtol, what):
    for k, ov in ora_states.items():
        dv = dev_states[k]
        assert dv.shape == ov.shape and dv.dtype == ov.dtype, (what, k, dv.dtype, ov.dtype)
        if ov.dtype == np.float32:
            assert np.allclose(dv, ov, rtol=rtol, atol=1e-6, equal_nan=True), (what, k, np.nanmax(np.abs(dv - ov)))
        else:
            assert np.array_equal(dv, ov), (what, k)


@pytest.mark.parametrize("nh,nf", [(1000, 50), (4099, 37)])
def test_economy_init_matches_oracle(mode, nh, nf):
    m = economy.create_economy_model(nh, nf, config=jx.ModelConfig(seed=7, rng_mode=mode))
    m.initialize()
    o = oeco.create_economy_model(nh, nf, config=ort.ModelConfig(seed=7, rng_mode=mode))
    o.initialize()
    _cmp_states(m.agent_collections["households"].states, o.agent_collections["households"].states, 2e-6, "hh")
    _cmp_states(m.agent_collections["consumer_firms"].states, o.agent_collections["consumer_firms"].states, 2e-6, "cf")


@pytest.mark.parametrize("nh,nf", [(2000, 50), (5003, 101)])
def test_economy_trajectory(mode, nh, nf):
    steps = 3
    m = economy.create_economy_model(nh, nf, config=jx.ModelConfig(seed=42, rng_mode=mode))
    r = m.run(steps=steps)
    o = oeco.create_economy_model(nh, nf, config=ort.ModelConfig(seed=42, rng_mode=mode))
    ro = o.run(steps=steps)
    assert list(r.keys()) == list(ro.keys())
    for k in oeco.METRIC_NAMES:
        a, b = np.array(r[k], dtype=np.float64), np.array(ro[k], dtype=np.float64)
        assert np.allclose(a, b, rtol=RTOL, atol=ATOL_DIFF.get(k, 1e-6)), (k, a, b)
    _cmp_states(m.agent_collections["households"].states, o.agent_collections["households"].states, 1e-4, "hh")
    _cmp_states(m.agent_collections["consumer_firms"].states, o.agent_collections["consumer_firms"].states, 1e-4, "cf")
    # integer / boolean state is bit-exact (employment transitions come from exact uniforms)
    assert np.array_equal(m.agent_collections["households"].states["employed"],
                          o.agent_collections["households"].states["employed"])


def test_economy_nan_regime_and_env_quirks(mode):
    """After the model has gone NaN the metrics are the nan_to_num defaults on both sides, and the
    env entries kept by update_environment's final dict comprehension never move."""
    m = economy.create_economy_model(1500, 40, config=jx.ModelConfig(seed=3, rng_mode=mode))
    r = m.run(steps=30)
    o = oeco.create_economy_model(1500, 40, config=ort.ModelConfig(seed=3, rng_mode=mode))
    ro = o.run(steps=30)
    for k in ("gdp", "wage_rate", "interest_rate", "inequality", "economic_health"):
        assert float(r[k][-1]) == pytest.approx(float(ro[k][-1]), rel=1e-6), k
    for k in ("inflation", "goods_availability", "consumer_price", "utility", "income_per_capita", "debt_to_gdp"):
        assert len(set(float(v) for v in r[k])) == 1, k          # frozen by :1731-1735
        assert float(r[k][0]) == pytest.approx(float(ro[k][0]), rel=1e-6)
    assert m._env_state["time_step"] == 30


def test_economy_gini_matches_sorted_formula(mode):
    """Histogram-rank Gini vs the reference's sorted formula on the device's own incomes."""
    m = economy.create_economy_model(100_003, 200, config=jx.ModelConfig(seed=5, rng_mode=mode))
    r = m.run(steps=2)
    inc = m.agent_collections["households"].states["income"]
    assert float(r["inequality"][-1]) == pytest.approx(float(oeco.gini_sorted(inc)), rel=1e-5)


def test_economy_households_only(mode):
    m = economy.create_economy_model(3000, 0, config=jx.ModelConfig(seed=9, rng_mode=mode))
    r = m.run(steps=4)
    o = oeco.create_economy_model(3000, 0, config=ort.ModelConfig(seed=9, rng_mode=mode))
    ro = o.run(steps=4)
    for k in ("gdp", "wage_rate", "interest_rate", "unemployment", "inequality", "economic_health"):
        assert np.allclose(np.array(r[k], dtype=np.float64), np.array(ro[k], dtype=np.float64), rtol=RTOL,
                           atol=ATOL_DIFF.get(k, 1e-6)), k


def test_economy_unregistered_parts_raise():
    with pytest.raises(jx.UnregisteredRuleError):
        economy.create_economy_model(100, 10, num_capital_firms=5)
    with pytest.raises(jx.UnregisteredRuleError):
        economy.create_economy_model(100, 10, enable_climate_module=True)


def test_economy_against_frozen_fixture(mode):
    """CUDA path vs the committed oracle fixture (tests/golden/oracle_golden.npz)."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.npz"))
    tag = "legacy" if mode == 0 else "part"
    m = economy.create_economy_model(1200, 30, config=jx.ModelConfig(seed=42, rng_mode=mode))
    m.initialize()
    assert np.allclose(m.agent_collections["households"].states["income"], gold[f"economy_{tag}_init_income"], rtol=2e-6)
    assert np.allclose(m.agent_collections["consumer_firms"].states["capital_stock"],
                       gold[f"economy_{tag}_init_capital"], rtol=2e-6)
    r = m.run(steps=3)
    for k in ("gdp", "wage_rate", "interest_rate", "unemployment", "inequality"):
        assert np.allclose([float(v) for v in r[k]], gold[f"economy_{tag}_{k}"][:3], rtol=RTOL, atol=1e-6), k


def test_economy_full_size_properties():
    """C4-B at bench size (49 M households + 1 M firms): employment flags stay boolean, every household
    was updated, the fused reductions equal host reductions of the downloaded columns, and the
    histogram-rank Gini equals the sorted formula evaluated on the host in float64."""
    nh, nf = 49_000_000, 1_000_000
    m = economy.create_economy_model(nh, nf, config=jx.ModelConfig(seed=42, rng_mode=1))
    r = m.run(steps=2)
    hh = m.agent_collections["households"].states
    emp = hh["employed"]
    assert emp.dtype == np.bool_ and emp.shape == (nh,)
    assert float(r["unemployment"][-1]) == pytest.approx(100.0 * (1.0 - emp.mean(dtype=np.float64)), rel=1e-5)
    inc = hh["income"].astype(np.float64)
    assert np.isfinite(inc).all() and (inc >= 0).all()
    x = np.sort(inc)
    n = x.shape[0]
    gini = 2.0 * np.dot(np.arange(1, n + 1, dtype=np.float64), x) / (n * x.sum()) - (n + 1) / n
    assert float(r["inequality"][-1]) == pytest.approx(gini, rel=2e-5)
    assert float(m._env_state["total_labor_supply"]) == pytest.approx(float(hh["labor_supply"].sum(dtype=np.float64)), rel=1e-6)
    cf = m.agent_collections["consumer_firms"].states
    assert (cf["age"] == 2).all()
