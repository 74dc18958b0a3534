"""CPU tests of the oracle: known-answer vectors, closed forms, the reference's behavioural
contract (ported from /root/reference/tests, which need JAX), frozen fixtures, C vs NumPy."""
import json
import os

import numpy as np
import pytest

from oracle import cfast, jaxlike as jl, rules as orules, runtime as ort

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _u32(x):
    return int(x, 16) if isinstance(x, str) else int(x)


@pytest.fixture(scope="module")
def kat():
    return json.load(open(os.path.join(GOLD, "threefry_kat.json")))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "oracle_golden.npz"))


# ---- published known-answer vectors -----------------------------------------------------------
def test_threefry_random123_kat(kat):
    for v in kat["threefry2x32"]:
        k, c, o = [_u32(x) for x in v["key"]], [_u32(x) for x in v["ctr"]], [_u32(x) for x in v["out"]]
        y0, y1 = jl.threefry2x32(k[0], k[1], c[0], c[1])
        assert [int(y0), int(y1)] == o


def test_jax_published_values(kat):
    assert jl.split(jl.PRNGKey(0), 2, jl.LEGACY).tolist() == kat["split_prngkey0_legacy"]
    assert jl.split(jl.PRNGKey(0), 2, jl.PARTITIONABLE).tolist() == kat["split_prngkey0_partitionable"]
    assert float(jl.uniform(jl.PRNGKey(0), (), mode=jl.LEGACY)) == pytest.approx(kat["uniform_prngkey0_legacy"], abs=1e-8)
    assert int(jl.random_bits(jl.PRNGKey(0), (), jl.PARTITIONABLE)) == _u32(kat["bits_prngkey0_partitionable"])


def test_jax_documented_draws(kat):
    """Outputs printed in JAX's documentation, float32 to the printed digits: they pin the bits -> uniform
    mapping of both stream layouts and the whole normal path (uniform on (-1, 1) + XLA's float32 erf_inv)."""
    f32 = lambda v: float(np.float32(v))
    assert f32(jl.normal(jl.PRNGKey(0), (1,), jl.LEGACY)[0]) == f32(kat["normal_prngkey0_shape1_legacy_docs"])
    sub = jl.split(jl.PRNGKey(0), 2, jl.LEGACY)[1]
    assert f32(jl.normal(sub, (1,), jl.LEGACY)[0]) == f32(kat["normal_after_first_split_of_prngkey0_shape1_legacy_docs"])
    assert f32(jl.normal(jl.PRNGKey(42), (), jl.LEGACY)) == f32(kat["normal_prngkey42_legacy_docs"])
    assert f32(jl.uniform(jl.PRNGKey(0), (), mode=jl.PARTITIONABLE)) == pytest.approx(kat["uniform_prngkey0_partitionable_docs"], abs=5e-7)
    assert f32(jl.normal(jl.PRNGKey(42), (), jl.PARTITIONABLE)) == f32(kat["normal_prngkey42_partitionable_docs"])
    # the tutorial's chained draws: split -> draw from the subkey -> carry the new key (the schedule of model.py:156)
    assert jl.split(jl.PRNGKey(42), 2, jl.LEGACY).tolist() == kat["split_prngkey42_legacy_docs"]
    for mode, name in ((jl.LEGACY, "legacy"), (jl.PARTITIONABLE, "partitionable")):
        key, got = jl.PRNGKey(42), []
        for _ in range(3):
            key, sub = jl.split(key, 2, mode)
            got.append(float(jl.normal(sub, (), mode)))
        assert got == kat[f"tutorial_chained_normal_draws_prngkey42_{name}_docs"]


def test_prng_structure(mode):
    key = jl.PRNGKey(7)
    # batched helpers agree with the scalar definitions
    ks = jl.split(key, 6, mode)
    assert np.array_equal(jl.split_batched(ks, 3, mode), np.stack([jl.split(k, 3, mode) for k in ks]))
    assert np.array_equal(jl.random_bits_scalar_batched(ks, mode),
                          np.array([jl.random_bits(k, (), mode) for k in ks], dtype=np.uint32))
    u = jl.uniform(key, (4096,), mode=mode)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0 and 0.45 < u.mean() < 0.55
    r = jl.randint(key, (4096,), 3, 17, mode)
    assert r.dtype == np.int32 and r.min() == 3 and r.max() == 16
    n = jl.normal(key, (20000,), mode)
    assert abs(float(n.mean())) < 0.03 and abs(float(n.std()) - 1.0) < 0.03
    p = jl.permutation(key, np.arange(100), mode)
    assert sorted(p.tolist()) == list(range(100)) and p.tolist() != list(range(100))


def test_feistel_is_a_bijection():
    for n in (1, 2, 3, 5, 17, 1000, 4097):
        p = jl.feistel_permute(np.arange(n), n, [11, 22, 33, 44])
        assert sorted(p.tolist()) == list(range(n))


# ---- frozen fixtures ------------------------------------------------------------------------------
def test_oracle_matches_frozen_fixtures(gold, mode):
    tag = "legacy" if mode == 0 else "part"
    key = jl.PRNGKey(42)
    assert np.array_equal(jl.split(key, 5, mode), gold[f"prng_{tag}_split5"])
    assert np.array_equal(jl.uniform(key, (9,), -1.0, 3.0, mode), gold[f"prng_{tag}_uniform9"])
    assert np.array_equal(jl.randint(key, (11,), 0, 1000000, mode), gold[f"prng_{tag}_randint11"])
    mk = orules.create_economy_model(num_consumers=2000, num_producers=500, config=ort.ModelConfig(seed=42, rng_mode=mode))
    r = mk.run(steps=40)
    np.testing.assert_allclose([float(v) for v in r["gdp"]], gold[f"market_{tag}_gdp"], rtol=1e-6)
    sc = orules.create_schelling_model(48, 1800, seed=5, config=ort.ModelConfig(seed=5, rng_mode=mode))
    sc.run(steps=15)
    assert np.array_equal(sc.agent_collections["agents"].states["position"], gold[f"schelling_{tag}_position"])


# ---- closed forms ------------------------------------------------------------------------------------
def test_random_walk_closed_form(mode):
    """All walkers share one deterministic trajectory: start 0.5, v 0.01, reflect at the bounds."""
    r = orules.RandomWalkModelNamed({"n_agents": 64, "steps": 120, "seed": 1}, rng_mode=mode).run()
    md, mx = np.array(r["mean_distance"], dtype=np.float64), np.array(r["max_distance"], dtype=np.float64)
    np.testing.assert_allclose(md, mx, rtol=1e-6)
    x, v, want = np.float32(0.5), np.float32(0.01), []
    for _ in range(120):
        nx = np.float32(x + v)
        if nx <= 0.0 or nx >= 1.0:
            v = np.float32(-v)
        x = np.float32(min(max(nx, np.float32(0.0)), np.float32(1.0)))
        d = np.float32(x - np.float32(0.5))
        want.append(np.sqrt(np.float32(d * d + d * d)))
    np.testing.assert_allclose(mx, np.array(want, dtype=np.float64), rtol=1e-6)
    # the unnamed variant reproduces the reference's quirk: 'walkers' is never found
    q = orules.RandomWalkModel({"n_agents": 64, "steps": 5, "seed": 1}, rng_mode=mode).run()
    assert q["mean_distance"] == [0.0] * 5 and q["time"] == [0] * 5 and q["num_red"] == [64] * 5


def test_growth_and_counter_closed_forms(mode):
    m = orules.create_test_model(growth_rate=0.1, initial_value=1.0, num_agents=10,
                                 config=ort.ModelConfig(seed=0, rng_mode=mode))
    r = m.run(steps=10)
    v = np.float32(1.0)
    for _ in range(10):
        v = np.float32(v * np.float32(1.1))
    assert r["avg_value"][-1] == pytest.approx(float(v), rel=1e-6)
    p = 1.0
    for _ in range(10):
        p += 0.1 * (1.2 - p)
    assert r["price_level"][-1] == p
    om = ort.Model(params={}, config=ort.ModelConfig(seed=0, rng_mode=mode),
                   update_state_fn=orules.counter_update_fn, metrics_fn=orules.counter_metrics_fn)
    om.add_agent_collection("consumers", ort.AgentCollection(orules.IncrementAgent(), 10))
    om.add_env_state("counter", 0)
    om.add_env_state("increment", 2.0)
    om.initialize()
    s0 = float(np.sum(om.agent_collections["consumers"].states["value"]))
    rr = om.run(steps=5)                                           # test_model.py:226-230
    assert rr["step_counter"] == [1, 2, 3, 4, 5]
    assert float(rr["total_value"][-1]) == pytest.approx(s0 + 10 * 5 * 2.0, rel=1e-5)


# ---- the reference's behavioural contract (tests/unit/test_model.py, test_agent.py) -------------------
def test_model_contract():
    with pytest.raises(ValueError):
        ort.AgentCollection(orules.IncrementAgent(), 0)                  # agent.py:83-84
    m = ort.Model()
    with pytest.raises(ValueError):
        m.initialize()                                                   # model.py:125-126
    with pytest.raises(RuntimeError):
        m.step()                                                         # model.py:152-153
    m = ort.Model(update_state_fn=orules.counter_update_fn, metrics_fn=orules.counter_metrics_fn,
                  config=ort.ModelConfig(steps=7))
    c = ort.AgentCollection(orules.IncrementAgent(), 10)
    with pytest.raises(ValueError):
        c.update({}, jl.PRNGKey(0), ort.ModelConfig())                   # agent.py:150-151
    with pytest.raises(TypeError):
        c.init(jl.PRNGKey(0), object())                                  # agent.py:103-104
    m.add_agent_collection("consumers", c)
    m.add_env_state("counter", 0)
    r = m.run()
    assert len(r["step"]) == 7 and r["step"][-1] == 7
    with pytest.raises(RuntimeError):
        m.add_agent_collection("late", c)                                # model.py:71-72
    r2 = m.run()                                                         # test_model.py:274-276
    assert len(r2["step"]) == 7 and r2["step"][0] == 8                   # time keeps counting, history resets
    assert m.state["env"]["counter"] == 0                                # _state['env'] is the init-time copy
    m2 = ort.Model(config=ort.ModelConfig(track_history=False), metrics_fn=orules.counter_metrics_fn)
    m2.add_agent_collection("consumers", ort.AgentCollection(orules.IncrementAgent(), 3))
    assert m2.run(steps=3) == {}
    m3 = ort.Model(config=ort.ModelConfig(collect_interval=4), metrics_fn=orules.counter_metrics_fn)
    m3.add_agent_collection("consumers", ort.AgentCollection(orules.IncrementAgent(), 3))
    assert m3.run(steps=10)["step"] == [4, 8]


def test_agent_collection_contract(mode):
    cfg = ort.ModelConfig(rng_mode=mode)
    c = ort.AgentCollection(orules.WealthAgent(), 10)
    c.init(jl.PRNGKey(0), cfg)
    w0, p = c.states["wealth"].copy(), c.states["productivity"]
    assert w0.shape == (10,) and (w0 >= 0).all() and (w0 <= 100).all() and (p >= 0.5).all() and (p <= 1.5).all()
    c.update({}, jl.PRNGKey(1), cfg)
    np.testing.assert_allclose(c.states["wealth"] - w0, p, rtol=1e-5)    # test_agent.py:204-209
    f = c.filter(lambda s: s["wealth"] > 50)
    assert (f.states["wealth"] > 50).all()
    with pytest.raises(ValueError):
        c.aggregate("nope")


def test_market_direction(mode):
    """test_integration.py:338-348: higher propensity / productivity => higher final GDP."""
    def gdp(**params):
        return float(orules.create_economy_model(params=params, config=ort.ModelConfig(seed=42, rng_mode=mode)).run()["gdp"][-1])
    assert gdp(propensity_to_consume=0.9) > gdp(propensity_to_consume=0.6)
    assert gdp(productivity=1.5) > gdp(productivity=0.8)


# ---- builder-authored rules: invariants ------------------------------------------------------------------
@pytest.mark.parametrize("periodic", [False, True])
def test_schelling_invariants(mode, periodic):
    om = orules.create_schelling_model(24, 400, periodic=periodic, seed=3, config=ort.ModelConfig(seed=3, rng_mode=mode))
    r = om.run(steps=10)
    st = om.agent_collections["agents"].states
    cells = st["position"][:, 0] * 24 + st["position"][:, 1]
    assert len(set(cells.tolist())) == 400                               # one agent per cell
    grid = om._env_state["grid"]
    assert (grid >= 0).sum() == 400 and np.array_equal(grid[st["position"][:, 0], st["position"][:, 1]], st["type"])
    ec = om._env_state["empty_cells"]
    assert sorted((ec[:, 0] * 24 + ec[:, 1]).tolist()) == np.nonzero(grid.reshape(-1) < 0)[0].tolist()
    assert int(r["total_moves"][-1]) == int(st["moves"].sum())
    assert r["percent_satisfied"][-1] >= r["percent_satisfied"][0]


def test_sir_invariants(mode):
    from jaxabm_b200.synthetic import ring_lattice_edges
    n = 500
    om = orules.create_sir_model(n, ring_lattice_edges(n, 2), beta=0.3, gamma=0.2, initial_infected=0.05,
                                 config=ort.ModelConfig(seed=1, rng_mode=mode))
    r = om.run(steps=30)
    tot = np.array(r["count_S"]) + np.array(r["count_I"]) + np.array(r["count_R"])
    assert (tot == n).all()
    assert (np.diff(np.array(r["count_S"])) <= 0).all() and (np.diff(np.array(r["count_R"])) >= 0).all()


# ---- C / OpenMP restatement against the NumPy oracle --------------------------------------------------------
@pytest.mark.parametrize("g,n,periodic", [(64, 3100, False), (37, 1000, True), (20, 300, False)])
def test_c_oracle_schelling(mode, g, n, periodic):
    om = orules.create_schelling_model(g, n, periodic=periodic, seed=11, config=ort.ModelConfig(seed=11, rng_mode=mode))
    types, pos = orules.schelling_initial_layout(g, n, 0.5, 11)
    f = cfast.SchellingFast(g, types, pos, 0.5, periodic, seed=11, mode=mode)
    r, fr = om.run(steps=10), f.run(10)
    st = om.agent_collections["agents"].states
    assert np.array_equal(st["position"], f.pos) and np.array_equal(st["moves"], f.moves)
    assert np.array_equal(st["satisfied"], f.satisfied.astype(bool))
    assert [int(v) for v in r["total_moves"]] == [int(v) for v in fr["total_moves"]]
    np.testing.assert_allclose(r["segregation_index"], fr["segregation_index"], rtol=1e-6)
    ec = om._env_state["empty_cells"]
    assert np.array_equal(ec[:, 0] * g + ec[:, 1], f.E)


def test_c_oracle_key_schedule(mode):
    a = ort.key_schedule(5, 2, True, 6, mode)
    b = cfast.key_schedule(5, 2, True, 6, mode)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_c_oracle_sir_and_walk_match_the_numpy_oracle(mode):
    """oracle/c: orc_sir_step / orc_walk_step (bench.py's cpu_baseline for --workload sir / walk) against the
    NumPy restatement: integer states exact, float32 walker columns exact, mean distance to summation order."""
    from jaxabm_b200 import synthetic
    n = 2500
    edges = synthetic.scale_free_edges(n, 3, 5)
    om = orules.create_sir_model(n, edges, beta=0.1, gamma=0.1, initial_infected=0.02, seed=2,
                                config=ort.ModelConfig(seed=2, rng_mode=mode))
    r = om.run(steps=12)
    f = cfast.SirFast(n, edges, beta=0.1, gamma=0.1, initial_infected=0.02, seed=2, mode=mode)
    fr = f.run(12)
    for k in ("count_S", "count_I", "count_R"):
        assert [int(v) for v in r[k]] == [int(v) for v in fr[k]]
    assert np.array_equal(f.state, np.asarray(om.agent_collections["agents"].states["state"]))
    # walkers: the scaled variant's keyed start, 40 steps (many bounces at |v| <= 0.01 near the walls)
    w = orules.ScaledRandomWalker()
    keys = jl.split(jl.PRNGKey(9), 4000, mode)
    s = w.init_batch(ort.ModelConfig(seed=9, rng_mode=mode), keys)
    s = {"position": s["position"], "velocity": s["velocity"] * np.float32(8.0),
         "color": np.zeros(4000, np.int32), "steps_taken": np.zeros(4000, np.int32)}
    fw = cfast.WalkFast(s["position"], s["velocity"])
    out = fw.run(40)
    for _ in range(40):
        s = w.step_batch(s, {"env": {"bounds": np.array([0.0, 1.0], dtype=np.float32)}})
    assert np.array_equal(fw.pos, s["position"]) and np.array_equal(fw.vel, s["velocity"])
    assert np.array_equal(fw.color, s["color"]) and np.array_equal(fw.steps_taken, s["steps_taken"]) and fw.color.sum() > 100
    d = orules.walker_distances(s["position"])
    assert float(out["max_distance"][-1]) == float(d.max())
    assert float(out["mean_distance"][-1]) == pytest.approx(float(d.mean(dtype=np.float64)), rel=1e-6)
