"""GPU tests at BASELINE.json's full sizes (size-independent properties + the C/OpenMP oracle
where it finishes in seconds) and against the frozen fixtures in tests/golden/."""
import os

import numpy as np
import pytest

import jaxabm_b200 as jx
from jaxabm_b200 import synthetic
from jaxabm_b200.rules import growth, market, random_walk, schelling, sir

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.npz")


def series(d, k):
    return np.array([float(v) for v in d[k]], dtype=np.float64)


def test_against_frozen_fixtures(mode):
    g = np.load(GOLD)
    tag = "legacy" if mode == 0 else "part"
    res = random_walk.RandomWalkModel({"n_agents": 1000, "steps": 100, "seed": 42, "name": "walkers", "rng_mode": mode}).run()
    assert np.array_equal(series(res._data, "max_distance"), g[f"walk_{tag}_max_distance"])
    np.testing.assert_allclose(series(res._data, "mean_distance"), g[f"walk_{tag}_mean_distance"], rtol=1e-6)
    mk = market.create_economy_model(num_consumers=2000, num_producers=500, config=jx.ModelConfig(seed=42, rng_mode=mode))
    r = mk.run(steps=40)
    for k in ("gdp", "price_level", "unemployment", "avg_utility", "avg_profit"):
        np.testing.assert_allclose(series(r, k), g[f"market_{tag}_{k}"], rtol=1e-5, atol=1e-7, err_msg=k)
    assert np.array_equal(mk.agent_collections["consumers"].states["income"], g[f"market_{tag}_income"])
    sc = schelling.create_schelling_model(48, 1800, seed=5, config=jx.ModelConfig(seed=5, rng_mode=mode))
    r = sc.run(steps=15)
    st = sc.agent_collections["agents"].states
    assert np.array_equal(st["position"], g[f"schelling_{tag}_position"])
    assert np.array_equal(st["moves"], g[f"schelling_{tag}_moves"])
    assert np.array_equal(st["satisfied"], g[f"schelling_{tag}_satisfied"])
    assert np.array_equal(series(r, "total_moves"), g[f"schelling_{tag}_total_moves"])
    sr = sir.create_sir_model(3000, synthetic.ring_lattice_edges(3000, 2), beta=0.3, gamma=0.1, initial_infected=0.02,
                              seed=9, config=jx.ModelConfig(seed=9, rng_mode=mode))
    r = sr.run(steps=25)
    for k in ("count_S", "count_I", "count_R"):
        assert np.array_equal(series(r, k), g[f"sir_{tag}_{k}"]), k
    assert np.array_equal(sr.agent_collections["agents"].states["state"], g[f"sir_{tag}_state"])
    gm = growth.create_test_model(growth_rate=0.07, adjustment_rate=0.13, initial_value=2.5, num_agents=64,
                                  config=jx.ModelConfig(seed=0, rng_mode=mode))
    r = gm.run(steps=30)
    assert np.array_equal(series(r, "price_level"), g[f"growth_{tag}_price_level"])
    np.testing.assert_allclose(series(r, "avg_value"), g[f"growth_{tag}_avg_value"], rtol=1e-6)


def test_schelling_full_size_vs_c_oracle():
    """C2 at full size (4096^2, 13 M agents): bit-exact against the C/OpenMP oracle over the first 64
    steps (~3.7 M agents move in step 1), then conservation invariants after 60 more."""
    from oracle import cfast
    G, N = 4096, 13_000_000
    types, pos = schelling.initial_layout(G, N, 0.5, 42)
    m = schelling.create_schelling_model(G, N, seed=42, types=types, positions=pos, config=jx.ModelConfig(seed=42, rng_mode=1))
    f = cfast.SchellingFast(G, types, pos, seed=42, mode=1)
    # 4 + 60 steps in two run() calls (state and _time_step persist), every metric row and the whole state
    # compared bit for bit after each: the active phase (millions of movers per step) and its decay
    for steps in (4, 60):
        r, fr = m.run(steps=steps), f.run(steps)
        assert [int(v) for v in r["total_moves"]] == [int(v) for v in fr["total_moves"]]
        assert np.array_equal(series(r, "percent_satisfied"), series(fr, "percent_satisfied"))
        np.testing.assert_allclose(series(r, "segregation_index"), series(fr, "segregation_index"), rtol=1e-6)
        st = m.agent_collections["agents"].states
        assert np.array_equal(st["position"], f.pos)
        assert np.array_equal(st["moves"], f.moves)
        assert np.array_equal(st["satisfied"], f.satisfied.astype(bool))
        assert np.array_equal(m._dev.download_grid().reshape(-1), f.grid)
        ec = m._dev.download_empty_cells()
        assert np.array_equal(ec[:, 0].astype(np.int64) * G + ec[:, 1], f.E)
    r2 = m.run(steps=60)
    st = m.agent_collections["agents"].states
    p = st["position"].astype(np.int64)
    cells = p[:, 0] * G + p[:, 1]
    assert np.unique(cells).size == N                                  # still one agent per cell
    grid = m._dev.download_grid()
    assert (grid >= 0).sum() == N and np.array_equal(grid.reshape(-1)[cells], st["type"])
    ec = m._dev.download_empty_cells().astype(np.int64)
    assert np.array_equal(np.sort(ec[:, 0] * G + ec[:, 1]), np.nonzero(grid.reshape(-1) < 0)[0])
    assert int(r2["total_moves"][-1]) == int(st["moves"].astype(np.int64).sum())
    ps = series(r2, "percent_satisfied")
    assert ps[-1] > 0.99 and ps[-1] == np.float32(st["satisfied"].mean())


@pytest.mark.parametrize("g,n,periodic,thr", [(1024, 800_000, False, 0.5), (1024, 800_000, True, 0.5),
                                              (2048, 3_000_000, True, 0.375), (1024, 1_000_000, False, 0.7),
                                              (1024, 1024 * 1024 - 5, False, 0.5)])
def test_schelling_bit_sliced_vs_c_oracle(mode, g, n, periodic, thr):
    """Row lengths that are multiples of 1024 take the bit-sliced kernel (csrc/schelling_bits.cuh: 1, 2
    strips per row here, 4 in the full-size test above): bit-exact against the C/OpenMP oracle over the
    active phase, periodic and non-periodic, several thresholds, nearly full grid."""
    from oracle import cfast
    types, pos = schelling.initial_layout(g, n, 0.5, 7)
    m = schelling.create_schelling_model(g, n, seed=7, types=types, positions=pos, periodic=periodic,
                                         similarity_threshold=thr, config=jx.ModelConfig(seed=7, rng_mode=mode))
    f = cfast.SchellingFast(g, types, pos, similarity_threshold=thr, periodic=periodic, seed=7, mode=mode)
    for steps in (1, 6, 9):
        r, fr = m.run(steps=steps), f.run(steps)
        assert [int(v) for v in r["total_moves"]] == [int(v) for v in fr["total_moves"]], steps
        assert np.array_equal(series(r, "percent_satisfied"), series(fr, "percent_satisfied"))
        np.testing.assert_allclose(series(r, "segregation_index"), series(fr, "segregation_index"), rtol=1e-6)
    st = m.agent_collections["agents"].states
    assert np.array_equal(st["position"], f.pos)
    assert np.array_equal(st["moves"], f.moves)
    assert np.array_equal(st["satisfied"], f.satisfied.astype(bool))
    assert np.array_equal(m._dev.download_grid().reshape(-1), f.grid)
    ec = m._dev.download_empty_cells()
    assert np.array_equal(ec[:, 0].astype(np.int64) * g + ec[:, 1], f.E)


def test_market_full_size_properties():
    """C4-A at 45 M + 5 M agents: totals reported by the fused reductions equal host-side sums of the
    downloaded columns; savings identity holds per agent."""
    nc, npr = 45_000_000, 5_000_000
    m = market.create_economy_model(num_consumers=nc, num_producers=npr, config=jx.ModelConfig(seed=42, rng_mode=1))
    r = m.run(steps=3)
    env = m.model_state()["env"]
    cs, ps = m.agent_collections["consumers"].states, m.agent_collections["producers"].states
    cons, prod = cs["consumption"], ps["production"]
    assert float(env["total_consumption"]) == pytest.approx(float(cons.sum(dtype=np.float64)), rel=2e-6)
    assert float(env["total_production"]) == pytest.approx(float(prod.sum(dtype=np.float64)), rel=2e-6)
    assert float(r["avg_utility"][-1]) == pytest.approx(float(cs["utility"].mean(dtype=np.float64)), rel=2e-6)
    assert float(r["avg_profit"][-1]) == pytest.approx(float(ps["profit"].mean(dtype=np.float64)), rel=2e-6)
    inc = cs["income"]
    assert inc.min() >= 0.8 and inc.max() <= 1.2000001 and abs(float(inc.mean(dtype=np.float64)) - 1.0) < 1e-3
    assert float(r["gdp"][-1]) == pytest.approx(float(env["total_production"]) * float(env["price_level"]), rel=1e-6)


def test_walk_full_size_properties():
    """C1 scaled (2^26 walkers): positions stay inside the bounds, every walker stepped K times,
    the fused max equals the host max and the fused mean the host mean."""
    n = 1 << 26
    m = random_walk.create_scaled_walk_model(n, config=jx.ModelConfig(seed=42, rng_mode=1))
    r = m.run(steps=7)
    st = m.agent_collections["walkers"].states
    pos = st["position"]
    assert pos.min() >= 0.0 and pos.max() <= 1.0
    assert (st["steps_taken"] == 7).all()
    d = np.sqrt(((pos - np.float32(0.5)) ** 2).sum(axis=1, dtype=np.float32))
    assert float(r["max_distance"][-1]) == float(d.max())
    assert float(r["mean_distance"][-1]) == pytest.approx(float(d.mean(dtype=np.float64)), rel=2e-6)


def test_sir_large_graph_properties():
    """C3-shaped input (scale-free, 2 M nodes / 20 M adjacency entries): S+I+R conserved, S
    non-increasing, R non-decreasing, and the same seed reproduces the same trajectory."""
    n = 2_000_000
    edges = synthetic.scale_free_edges(n, 5, 42)
    out = []
    for _ in range(2):
        m = sir.create_sir_model(n, edges, beta=0.05, gamma=0.1, initial_infected=0.01, seed=42,
                                 config=jx.ModelConfig(seed=42, rng_mode=1))
        out.append(m.run(steps=30))
    r = out[0]
    S, I, R = series(r, "count_S"), series(r, "count_I"), series(r, "count_R")
    assert ((S + I + R) == n).all() and (np.diff(S) <= 0).all() and (np.diff(R) >= 0).all() and I.max() > I[0]
    for k in ("count_S", "count_I", "count_R"):
        assert np.array_equal(series(out[0], k), series(out[1], k))


def test_sir_large_graph_vs_c_oracle(mode):
    """C3-shaped input (scale-free, 2 M nodes / 20 M adjacency entries, hubs of > 2048 entries: the heavy-row and
    long-row paths of both directions) bit for bit against the C/OpenMP oracle (``oracle/c::orc_sir_step``, itself
    pinned on the NumPy restatement in tests/test_oracle.py)."""
    from oracle import cfast
    n, steps = 2_000_000, 25
    edges = synthetic.scale_free_edges(n, 5, 42)
    m = sir.create_sir_model(n, edges, beta=0.05, gamma=0.1, initial_infected=0.01, seed=42,
                             config=jx.ModelConfig(seed=42, rng_mode=mode))
    r = m.run(steps=steps)
    f = cfast.SirFast(n, edges, beta=0.05, gamma=0.1, initial_infected=0.01, seed=42, mode=mode)
    fr = f.run(steps)
    for k in ("count_S", "count_I", "count_R"):
        assert np.array_equal(series(r, k), series(fr, k)), k
    assert np.array_equal(np.asarray(m.agent_collections["agents"].states["state"]), f.state)
    assert np.bincount(edges[:, 0]).max() > 2048


def test_sir_full_size_vs_c_oracle():
    """C3 at ITS OWN size -- 10 M agents, ~100 M adjacency entries (BASELINE.json configs[2]): counts of every step and
    the final state bit for bit against the C/OpenMP oracle, in the default direction-optimising mode (push while few
    rows are infected, then pull over the susceptible rows) and in the full ballot-segmented pull.  At this size the
    heavy-row list, the row-block table and the 32-bit adjacency offsets are in a different regime than at 2 M."""
    import os
    from oracle import cfast
    n, steps = 10_000_000, 30
    edges = synthetic.scale_free_edges(n, 5, 42)
    assert 99_000_000 < edges.shape[0] < 101_000_000
    f = cfast.SirFast(n, edges, beta=0.05, gamma=0.1, initial_infected=0.01, seed=42, mode=1)
    fr = f.run(steps)
    I = series(fr, "count_I")
    assert I.max() > 20 * I[0]                      # the run covers the take-off, i.e. both directions in auto mode
    for sir_mode in (None, "pull"):
        if sir_mode:
            os.environ["JXB_SIR_MODE"] = sir_mode
        try:
            m = sir.create_sir_model(n, edges, beta=0.05, gamma=0.1, initial_infected=0.01, seed=42,
                                     config=jx.ModelConfig(seed=42, rng_mode=1))
            r = m.run(steps=steps)
        finally:
            os.environ.pop("JXB_SIR_MODE", None)
        for k in ("count_S", "count_I", "count_R"):
            assert np.array_equal(series(r, k), series(fr, k)), (sir_mode, k)
        assert np.array_equal(np.asarray(m.agent_collections["agents"].states["state"]), f.state), sir_mode
        del m
