"""CPU checks of the C4-B economy restatement (oracle/economy.py): frozen fixtures, the
reference's observable quirks, and the Gini formula."""
import os

import numpy as np
import pytest

from oracle import economy as oeco, jaxlike as jl, runtime as ort

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_economy_matches_frozen_fixture(gold, mode):
    tag = "legacy" if mode == 0 else "part"
    ec = oeco.create_economy_model(1200, 30, config=ort.ModelConfig(seed=42, rng_mode=mode))
    ec.initialize()
    np.testing.assert_array_equal(ec.agent_collections["households"].states["income"], gold[f"economy_{tag}_init_income"])
    np.testing.assert_array_equal(ec.agent_collections["consumer_firms"].states["capital_stock"],
                                  gold[f"economy_{tag}_init_capital"])
    r = ec.run(steps=8)
    for k in ("gdp", "wage_rate", "interest_rate", "unemployment", "inequality", "economic_health"):
        np.testing.assert_allclose([float(v) for v in r[k]], gold[f"economy_{tag}_{k}"], rtol=1e-6, err_msg=k)
    np.testing.assert_array_equal(ec.agent_collections["households"].states["employed"], gold[f"economy_{tag}_employed"])


def test_beta_construction_vectorised_equals_scalar(mode):
    keys = jl.split(jl.PRNGKey(5), 64, mode)
    a, b = oeco._beta52(keys, mode), oeco._beta52_fast(keys, mode)
    assert np.array_equal(a, b) and (a > 0).all() and (a < 1).all()
    big = oeco._beta52_fast(jl.split(jl.PRNGKey(6), 20000, mode), mode)
    assert abs(float(big.mean()) - 5.0 / 7.0) < 5e-3            # Beta(5,2) mean


def test_env_carry_over_quirk(mode):
    """update_environment's last dict comprehension (advanced_economic_model.py:1731-1735) keeps every
    pre-existing entry outside its exclusion list: those metrics never move, and total_income
    (absent at the start) freezes at its first value."""
    m = oeco.create_economy_model(800, 20, config=ort.ModelConfig(seed=1, rng_mode=mode))
    r = m.run(steps=4)
    for k in ("inflation", "goods_availability", "consumer_price", "utility", "income_per_capita", "debt_to_gdp",
              "labor_market_tightness"):
        assert len(set(float(v) for v in r[k])) == 1, k
    assert float(r["inflation"][0]) == pytest.approx(2.0) and float(r["debt_to_gdp"][0]) == pytest.approx(60.0)
    assert len(set(float(v) for v in r["wage_rate"])) > 1 and len(set(float(v) for v in r["gdp"])) > 1
    first_income = m._env_state["total_income"]
    m.run(steps=2)
    assert m._env_state["total_income"] == first_income
    assert m._env_state["time_step"] == 6


def test_reference_model_degenerates_to_nan_defaults(mode):
    """0/0 in the firm's energy demand once inventory covers the demand share (:433-457) poisons the
    sums; compute_metrics then reports its nan_to_num defaults."""
    m = oeco.create_economy_model(1500, 40, config=ort.ModelConfig(seed=3, rng_mode=mode))
    r = m.run(steps=30)
    assert float(r["gdp"][-1]) == pytest.approx(0.1) and float(r["wage_rate"][-1]) == 1.0
    assert float(r["economic_health"][-1]) == 50.0 and float(r["inequality"][-1]) == 0.0
    assert np.isfinite([float(v) for v in r["gdp"]]).all()


def test_gini_sorted_formula():
    x = np.array([1, 1, 1, 1], dtype=np.float32)
    assert float(oeco.gini_sorted(x)) == pytest.approx(0.0, abs=1e-6)
    x = np.array([0, 0, 0, 10], dtype=np.float32)
    assert float(oeco.gini_sorted(x)) == pytest.approx(0.75, abs=1e-6)
    assert float(oeco.gini_sorted(np.zeros(5, dtype=np.float32))) == 0.0
    assert float(oeco.gini_sorted(np.array([1.0, np.nan], dtype=np.float32))) == 0.0
