"""Parity of the CUDA path (through the C ABI) against the CPU oracle, same seeded inputs.

Bar: bit-exact for integer / byte / index state and RNG-driven transitions; float32
trajectories within the tolerance written next to each comparison.
"""
import numpy as np
import pytest

import jaxabm_b200 as jx
from jaxabm_b200 import synthetic
from jaxabm_b200.rules import contract, growth, market, random_walk, schelling, sir
from oracle import facade as ofacade, jaxlike as jl, rules as orules, runtime as ort

pytestmark = pytest.mark.gpu


def series(d, k):
    return np.array([float(v) for v in d[k]], dtype=np.float64)


# ---------------------------------------------------------------------------------- C1
@pytest.mark.parametrize("name", [None, "walkers"])
def test_random_walk_facade(mode, name):
    p = {"n_agents": 1000, "steps": 100, "seed": 42, "rng_mode": mode}
    if name:
        p["name"] = name
    model = random_walk.RandomWalkModel(p)
    res = model.run()
    oc = orules.RandomWalkModelNamed if name else orules.RandomWalkModel
    om = oc({"n_agents": 1000, "steps": 100, "seed": 42}, rng_mode=mode)
    ores = om.run()
    assert list(res._data["step"]) == list(ores["step"])
    for k in ("mean_x", "mean_y", "num_red", "num_blue", "time", "max_distance"):
        assert np.array_equal(series(res._data, k), series(ores, k)), k
    np.testing.assert_allclose(series(res._data, "mean_distance"), series(ores, "mean_distance"), rtol=1e-6)
    st, ost = model.walkers.collection.states, om.walkers.collection.states
    for k in ("position", "velocity", "color", "steps_taken"):
        assert np.array_equal(st[k], ost[k]), k
    assert not any(k.startswith("agents.") for k in res._data)        # F11


@pytest.mark.parametrize("n", [1, 3, 1000, 100003])
def test_scaled_walker(mode, n):
    m = random_walk.create_scaled_walk_model(n, seed=5, config=jx.ModelConfig(seed=5, rng_mode=mode))
    r = m.run(steps=25)
    om = ort.Model(params={}, config=ort.ModelConfig(seed=5, rng_mode=mode),
                   update_state_fn=lambda e, a, p, k: e,
                   metrics_fn=lambda e, a, p: {"mean_distance": np.float32(np.mean(orules.walker_distances(a["walkers"]["position"]))),
                                               "max_distance": np.float32(np.max(orules.walker_distances(a["walkers"]["position"])))})
    om.add_agent_collection("walkers", ort.AgentCollection(orules.ScaledRandomWalker(), n))
    om.add_env_state("bounds", np.array([0.0, 1.0], dtype=np.float32))
    orr = om.run(steps=25)
    st, ost = m.agent_collections["walkers"].states, om.agent_collections["walkers"].states
    for k in ("position", "velocity", "color", "steps_taken"):
        assert np.array_equal(st[k], ost[k]), k                          # elementwise fp32: exact
    assert np.array_equal(series(r, "max_distance"), series(orr, "max_distance"))
    np.testing.assert_allclose(series(r, "mean_distance"), series(orr, "mean_distance"), rtol=2e-6)


# ---------------------------------------------------------------------------------- C4-A
@pytest.mark.parametrize("nc,npr", [(20, 5), (4097, 1023), (200000, 50001)])
def test_market(mode, nc, npr):
    m = market.create_economy_model(num_consumers=nc, num_producers=npr, config=jx.ModelConfig(seed=42, rng_mode=mode))
    r = m.run(steps=40)
    om = orules.create_economy_model(num_consumers=nc, num_producers=npr, config=ort.ModelConfig(seed=42, rng_mode=mode))
    orr = om.run(steps=40)
    assert list(r.keys()) == list(orr.keys())
    for k in ("gdp", "price_level", "unemployment", "avg_utility", "avg_profit"):
        np.testing.assert_allclose(series(r, k), series(orr, k), rtol=1e-5, atol=1e-7, err_msg=k)
    cs, ocs = m.agent_collections["consumers"].states, om.agent_collections["consumers"].states
    assert np.array_equal(cs["income"], ocs["income"])                  # threefry-driven init: exact
    for k in ("savings", "consumption", "utility"):
        np.testing.assert_allclose(cs[k], ocs[k], rtol=2e-5, atol=1e-6, err_msg=k)
    ps, ops = m.agent_collections["producers"].states, om.agent_collections["producers"].states
    for k in ("capital", "production", "profit"):
        np.testing.assert_allclose(ps[k], ops[k], rtol=2e-5, atol=1e-6, err_msg=k)
    # env written by the device tail is visible through model_state()
    assert float(m.model_state()["env"]["price_level"]) == pytest.approx(float(orr["price_level"][-1]), rel=1e-5)


def test_market_consumers_only(mode):
    m = market.create_economy_model(num_consumers=50, num_producers=0, config=jx.ModelConfig(seed=1, rng_mode=mode))
    r = m.run(steps=5)
    assert "avg_profit" not in r and "avg_utility" in r


# ---------------------------------------------------------------------------------- C5 model
def test_growth_model(mode):
    m = growth.create_test_model(growth_rate=0.07, adjustment_rate=0.13, initial_value=2.5, num_agents=1001,
                                 config=jx.ModelConfig(seed=0, rng_mode=mode))
    r = m.run(steps=30)
    orr = orules.create_test_model(growth_rate=0.07, adjustment_rate=0.13, initial_value=2.5, num_agents=1001,
                                   config=ort.ModelConfig(seed=0, rng_mode=mode)).run(steps=30)
    assert np.array_equal(series(r, "price_level"), series(orr, "price_level"))     # float64 recursion: exact
    assert np.array_equal(series(r, "price_gap"), series(orr, "price_gap"))
    np.testing.assert_allclose(series(r, "avg_value"), series(orr, "avg_value"), rtol=1e-6)
    # closed form v0 * fl32(1+g)^T
    v = np.float32(2.5)
    for _ in range(30):
        v = np.float32(v * np.float32(1.07))
    assert np.all(m.agent_collections["consumers"].states["value"] == v)


def test_counter_model(mode):
    m = jx.JaxModel(params={}, config=jx.ModelConfig(seed=0, rng_mode=mode),
                    update_state_fn=contract.update_state_fn, metrics_fn=contract.metrics_fn)
    m.add_agent_collection("consumers", jx.AgentCollection(contract.IncrementAgent(), 10))
    m.add_env_state("counter", 0)
    m.add_env_state("increment", 2.0)
    om = ort.Model(params={}, config=ort.ModelConfig(seed=0, rng_mode=mode),
                   update_state_fn=orules.counter_update_fn, metrics_fn=orules.counter_metrics_fn)
    om.add_agent_collection("consumers", ort.AgentCollection(orules.IncrementAgent(), 10))
    om.add_env_state("counter", 0)
    om.add_env_state("increment", 2.0)
    r, orr = m.run(steps=7), om.run(steps=7)
    assert [int(v) for v in r["step_counter"]] == [int(v) for v in orr["step_counter"]]
    np.testing.assert_allclose(series(r, "total_value"), series(orr, "total_value"), rtol=1e-6)
    assert np.array_equal(m.agent_collections["consumers"].states["value"],
                          om.agent_collections["consumers"].states["value"])


def test_wealth_collection(mode):
    key = jx.random.PRNGKey(0)
    cfg, ocfg = jx.ModelConfig(rng_mode=mode), ort.ModelConfig(rng_mode=mode)
    c = jx.AgentCollection(contract.WealthAgent(), 10)
    c.init(key, cfg)
    oc = ort.AgentCollection(orules.WealthAgent(), 10)
    oc.init(key, ocfg)
    for k in ("wealth", "productivity"):
        assert np.array_equal(c.states[k], oc.states[k]), k
    k2 = jx.random.split(key, 2, mode)[1]
    c.update({"wage_rate": 1.5}, k2, cfg)
    oc.update({"wage_rate": 1.5}, k2, ocfg)
    assert np.array_equal(c.states["wealth"], oc.states["wealth"])
    f = c.filter(lambda s: s["wealth"] > 50.0)
    assert np.all(f.states["wealth"] > 50.0) and f.num_agents == int(np.sum(oc.states["wealth"] > 50.0))
    assert float(c.aggregate("wealth")) == pytest.approx(float(np.mean(oc.states["wealth"])), rel=1e-6)


# ---------------------------------------------------------------------------------- C2
@pytest.mark.parametrize("g0,n,periodic", [(64, 3100, False), (64, 3100, True), (48, 2000, False),
                                            (37, 1000, False), (37, 1000, True), (16, 250, False),
                                            (128, 16000, False), (20, 300, False)])
def test_schelling(mode, g0, n, periodic):
    steps = 12
    m = schelling.create_schelling_model(g0, n, periodic=periodic, seed=11, config=jx.ModelConfig(seed=11, rng_mode=mode))
    om = orules.create_schelling_model(g0, n, periodic=periodic, seed=11, config=ort.ModelConfig(seed=11, rng_mode=mode))
    r, orr = m.run(steps=steps), om.run(steps=steps)
    st, ost = m.agent_collections["agents"].states, om.agent_collections["agents"].states
    for k in ("type", "position", "satisfied", "moves"):
        assert np.array_equal(st[k], ost[k]), k
    assert [int(v) for v in r["total_moves"]] == [int(v) for v in orr["total_moves"]]
    assert np.array_equal(series(r, "percent_satisfied"), series(orr, "percent_satisfied"))
    np.testing.assert_allclose(series(r, "segregation_index"), series(orr, "segregation_index"), rtol=1e-6)
    assert np.array_equal(m._dev.download_grid(), om._env_state["grid"])
    assert np.array_equal(m._dev.download_empty_cells(), om._env_state["empty_cells"])
    # continue the same models: state and key chain persist across run() calls (model.py:203)
    r2, orr2 = m.run(steps=3), om.run(steps=3)
    assert list(r2["step"]) == list(orr2["step"]) == [steps + 1, steps + 2, steps + 3]
    assert np.array_equal(m.agent_collections["agents"].states["position"],
                          om.agent_collections["agents"].states["position"])


def test_schelling_facade(mode):
    p = {"grid_size": 32, "n_agents": 800, "steps": 6, "seed": 3, "rng_mode": mode}
    model = schelling.SchellingModel(p)
    res = model.run()
    om = orules.create_schelling_model(32, 800, seed=3, config=ort.ModelConfig(seed=3, steps=6, rng_mode=mode))
    orr = om.run()
    assert [int(v) for v in res._data["total_moves"]] == [int(v) for v in orr["total_moves"]]
    assert np.array_equal(model.agents.position, om.agent_collections["agents"].states["position"])
    assert np.array_equal(model.grid_state, om._env_state["grid"])


def test_schelling_full_grid_no_moves(mode):
    # every cell occupied: nobody can move, everything else still evaluated
    g = 16
    m = schelling.create_schelling_model(g, g * g, seed=2, config=jx.ModelConfig(seed=2, rng_mode=mode))
    om = orules.create_schelling_model(g, g * g, seed=2, config=ort.ModelConfig(seed=2, rng_mode=mode))
    r, orr = m.run(steps=3), om.run(steps=3)
    assert [int(v) for v in r["total_moves"]] == [0, 0, 0] == [int(v) for v in orr["total_moves"]]
    assert np.array_equal(m.agent_collections["agents"].states["satisfied"],
                          om.agent_collections["agents"].states["satisfied"])


# ---------------------------------------------------------------------------------- C3
def _sir_pair(n, edges, mode, **kw):
    m = sir.create_sir_model(n, edges, seed=9, config=jx.ModelConfig(seed=9, rng_mode=mode), **kw)
    om = orules.create_sir_model(n, edges, seed=9, config=ort.ModelConfig(seed=9, rng_mode=mode), **kw)
    return m, om


@pytest.mark.parametrize("graph", ["ring", "scale_free", "star", "empty"])
def test_sir(mode, graph):
    if graph == "ring":
        n, edges = 5000, synthetic.ring_lattice_edges(5000, 3)
    elif graph == "scale_free":
        n, edges = 20011, synthetic.scale_free_edges(20011, 4, 1)
    elif graph == "star":       # one hub with degree >> tile size, leaves of degree 1
        n = 9000
        hub = np.zeros(n - 1, dtype=np.int32)
        leaves = np.arange(1, n, dtype=np.int32)
        edges = np.concatenate([np.stack([hub, leaves], 1), np.stack([leaves, hub], 1)])
    else:
        n, edges = 300, np.zeros((0, 2), dtype=np.int32)
    m, om = _sir_pair(n, edges, mode, beta=0.2, gamma=0.1, initial_infected=0.02)
    for chunk in (1, 7, 12):
        r, orr = m.run(steps=chunk), om.run(steps=chunk)
        for k in ("count_S", "count_I", "count_R"):
            assert [int(v) for v in r[k]] == [int(v) for v in orr[k]], (k, chunk)
        assert np.array_equal(m.agent_collections["agents"].states["state"],
                              om.agent_collections["agents"].states["state"])


# ---------------------------------------------------------------------------------- ensembles
def test_sensitivity_analysis_growth(mode):
    import functools
    ranges = {"growth_rate": (0.05, 0.2), "adjustment_rate": (0.05, 0.3)}
    metrics = ["avg_value", "price_level", "price_gap"]

    def fac(params=None, config=None):
        config.rng_mode = mode
        return growth.create_test_model(initial_value=1.0, num_agents=777, params=params, config=config)

    sa = jx.SensitivityAnalysis(fac, ranges, metrics, num_samples=16, seed=0)
    out = sa.run(verbose=False)
    for i in range(16):
        params = {p: float(sa.samples[i, j]) for j, p in enumerate(ranges)}
        om = orules.create_test_model(initial_value=1.0, num_agents=777, params=params,
                                      config=ort.ModelConfig(seed=i + 1000, rng_mode=mode))
        orr = om.run()
        np.testing.assert_allclose(out["avg_value"][i], orr["avg_value"][-1], rtol=2e-6)
        assert out["price_level"][i] == np.float32(orr["price_level"][-1])
        assert out["price_gap"][i] == orr["price_gap"][-1]
    idx = sa.sobol_indices()
    assert set(idx) == set(metrics)
    # the squared-correlation proxy (analysis.py:186-201) recomputed from the oracle's values
    vals = np.array(out["avg_value"], dtype=np.float32)
    vn = (vals - vals.mean()) / (vals.std() + np.float32(1e-8))
    pv = sa.samples[:, 0]
    pn = (pv - pv.mean()) / (pv.std() + np.float32(1e-8))
    assert idx["avg_value"]["growth_rate"] == pytest.approx(float(np.mean(pn * vn) ** 2), rel=1e-4)
    assert idx["avg_value"]["growth_rate"] > idx["avg_value"]["adjustment_rate"]


def test_ensemble_market_matches_single_runs(mode):
    from jaxabm_b200 import ensemble
    models, singles = [], []
    for i, ptc in enumerate([0.6, 0.7, 0.8, 0.9]):
        kw = dict(num_consumers=300, num_producers=70, params={"propensity_to_consume": ptc, "productivity": 1.0 + 0.1 * i})
        models.append(market.create_economy_model(config=jx.ModelConfig(seed=50 + i, rng_mode=mode), **kw))
        singles.append(market.create_economy_model(config=jx.ModelConfig(seed=50 + i, rng_mode=mode), **kw))
    last, _ = ensemble.run_last_metrics(models, steps=20)
    for i, s in enumerate(singles):
        r = s.run(steps=20)
        for k in ("gdp", "price_level", "avg_utility", "avg_profit"):
            np.testing.assert_allclose(float(last[k][i]), float(r[k][-1]), rtol=2e-6, err_msg=k)


def test_calibrator_evaluate_robust(mode):
    def fac(params=None, config=None):
        config.rng_mode = mode
        return market.create_economy_model(num_consumers=200, num_producers=40, params=params, config=config)

    cal = jx.CoreModelCalibrator(fac, {"propensity_to_consume": 0.7}, {"gdp": 60.0, "price_level": 1.0},
                                 method="es", max_iterations=2, evaluation_steps=10, seed=0)
    key = jl.PRNGKey(0)
    seeds = []
    for _ in range(3):
        key, sub = jl.split(key, 2)
        seeds.append(int(jl.randint(sub, (), 0, 1_000_000)))
    loss, ci = cal._evaluate_params_robust({"propensity_to_consume": 0.7})
    vals = {"gdp": [], "price_level": []}
    for s in seeds:
        orr = orules.create_economy_model(num_consumers=200, num_producers=40, params={"propensity_to_consume": 0.7},
                                          config=ort.ModelConfig(seed=s, rng_mode=mode)).run(steps=10)
        for k in vals:
            vals[k].append(float(orr[k][-1]))
    want = sum((abs(np.mean(np.array(v, dtype=np.float32)) - t) / (abs(t) + 1e-8)) ** 2
               for v, t in ((vals["gdp"], 60.0), (vals["price_level"], 1.0)))
    assert loss == pytest.approx(float(want), rel=1e-4)
    best = cal.calibrate(verbose=False)
    assert set(best) == {"propensity_to_consume"} and len(cal.loss_history) == 2


def test_saltelli_sobol_through_the_ensemble_kernel(mode):
    """True Sobol indices (SensitivityAnalysis.run_saltelli) on the growth model: avg_value depends on
    growth_rate only, price_gap on adjustment_rate only -- one ensemble launch for all (P+2)n runs."""
    from jaxabm_b200.analysis import SensitivityAnalysis
    from jaxabm_b200.rules import growth

    def factory(params=None, config=None):
        config.rng_mode = mode
        return growth.create_test_model(params=params, config=config, num_agents=256, initial_value=1.0)

    sa = SensitivityAnalysis(factory, {"growth_rate": (0.05, 0.2), "adjustment_rate": (0.05, 0.3)},
                             ["avg_value", "price_gap"], num_samples=8, seed=1)
    idx = sa.run_saltelli(num_base=256, steps=20)
    assert sa.last_device_seconds > 0.0
    assert idx["avg_value"]["ST"]["growth_rate"] > 0.9 and idx["avg_value"]["ST"]["adjustment_rate"] < 0.02
    assert idx["price_gap"]["ST"]["adjustment_rate"] > 0.9 and idx["price_gap"]["ST"]["growth_rate"] < 0.02


# ------------------------------------------------------------- C5: every shape of the ensemble kernel
ENS_SHAPES = {   # env overrides -> the launch shape they force (csrc/engine.cu::jxb_ensemble_run)
    "default": {},
    "cluster4": {"JXB_ENS_MIN_CLUSTER": "4"},
    "cluster8": {"JXB_ENS_MIN_CLUSTER": "8"},
    "l2_scratch": {"JXB_ENS_MAX_CLUSTER": "1"},
}


@pytest.mark.parametrize("shape", list(ENS_SHAPES))
def test_ensemble_c5_size_every_launch_shape(mode, shape, monkeypatch):
    """C5 at its own replica size (100 000 agents = 400 KB of state per replica): the default launch is a 2-CTA
    cluster whose CTAs exchange their partial rows through distributed shared memory; 4- and 8-CTA clusters and the
    per-CTA L2-scratch fallback are forced through the environment.  Every shape against the oracle's serial runs
    (analysis.py:113-157: seed i + 1000, last value per metric) over the configured 200 steps."""
    from jaxabm_b200 import ensemble
    for k, v in ENS_SHAPES[shape].items():
        monkeypatch.setenv(k, v)
    n, steps, R = 100_000, 200, 5
    rng = np.random.RandomState(3)
    g, a = rng.uniform(0.05, 0.2, R), rng.uniform(0.05, 0.3, R)
    models = [growth.create_test_model(params={"growth_rate": float(g[i]), "adjustment_rate": float(a[i])}, initial_value=1.0,
                                       config=jx.ModelConfig(seed=i + 1000, steps=steps, rng_mode=mode), num_agents=n)
              for i in range(R)]
    last, secs = ensemble.run_last_metrics(models, steps=steps)
    assert secs > 0
    for i in range(R):
        om = orules.create_test_model(params={"growth_rate": float(g[i]), "adjustment_rate": float(a[i])}, initial_value=1.0,
                                      config=ort.ModelConfig(seed=i + 1000, steps=steps, rng_mode=mode), num_agents=n)
        orr = om.run()
        assert float(orr["avg_value"][-1]) > 1000.0                     # (1 + g)^200: the values did evolve
        np.testing.assert_allclose(float(last["avg_value"][i]), float(orr["avg_value"][-1]), rtol=2e-6)
        assert np.float32(last["price_level"][i]) == np.float32(orr["price_level"][-1])
        assert last["price_gap"][i] == orr["price_gap"][-1]


@pytest.mark.parametrize("shape", ["default", "cluster2", "cluster8", "l2_scratch", "smem_forced_scratch"])
def test_ensemble_market_every_launch_shape(mode, shape, monkeypatch):
    """Two collections with per-agent keyed init (global indices across the cluster's slices) and four reduction
    slots: 30 000 consumers + 8 000 producers = 576 KB per replica (default: a 4-CTA cluster).  Every launch shape
    against single Model.run()s of the same replicas on the device, and those against the oracle."""
    from jaxabm_b200 import ensemble
    env = {"default": {}, "cluster2": {"JXB_ENS_MIN_CLUSTER": "2", "JXB_ENS_MAX_CLUSTER": "2"},
           "cluster8": {"JXB_ENS_MIN_CLUSTER": "8"}, "l2_scratch": {"JXB_ENS_MAX_CLUSTER": "1"},
           "smem_forced_scratch": {"JXB_ENS_FORCE_SCRATCH": "1"}}[shape]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    # cluster2 is forced on a state that would fit ONE CTA, smem_forced_scratch on one that would fit shared memory
    nc, npr = {"cluster2": (9_000, 2_000), "smem_forced_scratch": (3_000, 801)}.get(shape, (30_000, 8_000))
    kws = [dict(num_consumers=nc, num_producers=npr, params={"propensity_to_consume": ptc, "productivity": 1.0 + 0.1 * i})
           for i, ptc in enumerate([0.6, 0.75, 0.9])]
    models = [market.create_economy_model(config=jx.ModelConfig(seed=70 + i, rng_mode=mode), **kw) for i, kw in enumerate(kws)]
    last, _ = ensemble.run_last_metrics(models, steps=25)
    for i, kw in enumerate(kws):
        r = market.create_economy_model(config=jx.ModelConfig(seed=70 + i, rng_mode=mode), **kw).run(steps=25)
        orr = orules.create_economy_model(config=ort.ModelConfig(seed=70 + i, rng_mode=mode), **kw).run(steps=25)
        for k in ("gdp", "price_level", "avg_utility", "avg_profit"):
            np.testing.assert_allclose(float(last[k][i]), float(r[k][-1]), rtol=2e-6, err_msg=k)
            np.testing.assert_allclose(float(last[k][i]), float(orr[k][-1]), rtol=1e-5, err_msg=k)
