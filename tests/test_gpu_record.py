"""SURVEY.md 8 f4 on the device: per-agent time series (``agents.<name>.<var>``, jaxabm/agentpy.py:1103-1106) and
``AgentCollection.filter`` as a stream compaction (jaxabm/agent.py:213-243), against the CPU oracle."""
import numpy as np
import pytest

import jaxabm_b200 as jx
from jaxabm_b200 import synthetic
from jaxabm_b200.rules import market, random_walk, schelling, sir
from oracle import rules as orules, runtime as ort

pytestmark = pytest.mark.gpu


def _oracle_series(om, cname, var, steps, ci=1):
    """the oracle stepped one step at a time; the column after every recorded step"""
    out = []
    for t in range(1, steps + 1):
        om.step()
        if t % ci == 0:
            out.append(np.array(om.agent_collections[cname].states[var]))
    return np.stack(out) if out else None


@pytest.mark.parametrize("ci", [1, 3])
def test_series_market(mode, ci):
    """Graph-replayed step kernels: snapshots of two consumer columns and one producer column at every recorded
    step, over two run() calls (the ring is per run; _time_step persists)."""
    kw = dict(num_consumers=5003, num_producers=1001)
    m = market.create_economy_model(config=jx.ModelConfig(seed=4, rng_mode=mode, collect_interval=ci), **kw)
    m.record_agent_series("consumers", ["savings", "utility"])
    m.record_agent_series("producers", "capital")
    om = orules.create_economy_model(config=ort.ModelConfig(seed=4, rng_mode=mode, collect_interval=ci), **kw)
    om.initialize()
    t_done = 0
    for steps in (40, 7):
        r = m.run(steps=steps)
        n_rec = (t_done + steps) // ci - t_done // ci
        assert len(r["step"]) == n_rec
        want = {k: [] for k in (("consumers", "savings"), ("consumers", "utility"), ("producers", "capital"))}
        for t in range(t_done + 1, t_done + steps + 1):
            om.step()
            if t % ci == 0:
                for (c, v) in want:
                    want[(c, v)].append(np.array(om.agent_collections[c].states[v]))
        t_done += steps
        for (c, v), rows in want.items():
            got = m.agent_series[f"agents.{c}.{v}"]
            assert got.shape == (n_rec, kw["num_consumers"] if c == "consumers" else kw["num_producers"])
            np.testing.assert_allclose(got, np.stack(rows), rtol=1e-5, atol=1e-7, err_msg=f"{c}.{v}")
        if t_done % ci == 0:          # the last step was a recording step: the last snapshot is the live column
            assert np.array_equal(m.agent_series["agents.producers.capital"][-1], m.agent_collections["producers"].states["capital"])


def test_series_facade_results_keys(mode):
    """The facade keeps the reference's default (no agents.* keys) and fills them when the model opts in:
    vector-valued position f32[N,2] -> [T, N, 2]."""
    class Recorded(random_walk.RandomWalkModel):
        def setup(self):
            super().setup()
            self.record_agents("walkers", ["position", "color"])
    p = {"n_agents": 300, "steps": 25, "seed": 42, "name": "walkers", "rng_mode": mode}
    plain = random_walk.RandomWalkModel(dict(p)).run()
    assert not any(k.startswith("agents.") for k in plain._data)
    res = Recorded(dict(p)).run()
    pos, col = res._data["agents.walkers.position"], res._data["agents.walkers.color"]
    assert pos.shape == (25, 300, 2) and pos.dtype == np.float32 and col.shape == (25, 300) and col.dtype == np.int32
    om = orules.RandomWalkModelNamed({"n_agents": 300, "steps": 25, "seed": 42}, rng_mode=mode)
    om.run()
    assert np.array_equal(pos[-1], om.walkers.collection.states["position"])
    # walkers all start at (0.5, 0.5) with v = (0.01, 0.01): step t sits at 0.5 + 0.01 t until the first bounce
    np.testing.assert_allclose(pos[9, :, 0], np.float32(0.5) + 10 * np.float32(0.01), rtol=1e-6)
    series = res.variables.walkers.position                            # the Results attribute path of agentpy.py:618-806
    assert len(series) == 25 and np.array_equal(series[-1], pos[-1])


@pytest.mark.parametrize("bands", ["0", "1"])
@pytest.mark.parametrize("grid,n", [(64, 3100), (1024, 800_000)])
def test_series_schelling_persistent_kernel(mode, grid, n, bands, monkeypatch):
    """The persistent cooperative kernel is cut at the recording steps (bands = "1": the band kernels, whose snapshot
    launches sit between the steps of the captured graph): position / moves / the lazily materialised 'satisfied'
    column after every recorded step equal the oracle's, bit for bit; the run itself is unchanged."""
    monkeypatch.setenv("JXB_GRID_BANDS", bands)
    ci, steps = 2, 8
    kw = dict(seed=5, config=jx.ModelConfig(seed=9, rng_mode=mode, collect_interval=ci))
    m = schelling.create_schelling_model(grid, n, **kw)
    m.record_agent_series("agents", ["position", "satisfied", "moves"])
    r = m.run(steps=steps)
    ref = schelling.create_schelling_model(grid, n, **kw)
    r0 = ref.run(steps=steps)
    assert [int(v) for v in r["total_moves"]] == [int(v) for v in r0["total_moves"]]
    for k in ("position", "satisfied", "moves"):
        assert np.array_equal(m.agent_collections["agents"].states[k], ref.agent_collections["agents"].states[k]), k
        assert np.array_equal(m.agent_series[f"agents.agents.{k}"][-1], ref.agent_collections["agents"].states[k]), k
    if grid <= 64:
        om = orules.create_schelling_model(grid, n, seed=5, config=ort.ModelConfig(seed=9, rng_mode=mode, collect_interval=ci))
        rows = {k: [] for k in ("position", "satisfied", "moves")}
        for _ in range(steps // ci):              # run() in chunks: the seeded layout is placed by the first run()
            om.run(steps=ci)
            for k in rows:
                rows[k].append(np.array(om.agent_collections["agents"].states[k]))
        for k in rows:
            assert np.array_equal(m.agent_series[f"agents.agents.{k}"], np.stack(rows[k])), k


def test_series_sir_state(mode):
    """SIR keeps int8 state + a bitmap internally; the recorded API column int32[N] is unpacked inside the step graph
    (direction-optimising mode: push and pull steps both occur)."""
    n, steps = 40_000, 16
    edges = synthetic.scale_free_edges(n, 4, 7)
    kw = dict(beta=0.2, gamma=0.1, initial_infected=0.01, seed=3)
    m = sir.create_sir_model(n, edges, config=jx.ModelConfig(seed=3, rng_mode=mode), **kw)
    m.record_agent_series("agents", "state")
    r = m.run(steps=steps)
    om = orules.create_sir_model(n, edges, config=ort.ModelConfig(seed=3, rng_mode=mode), **kw)
    om.initialize()
    want = _oracle_series(om, "agents", "state", steps)
    got = m.agent_series["agents.agents.state"]
    assert got.dtype == np.int32 and np.array_equal(got, want)
    assert [int(v) for v in r["count_I"]] == [int((row == 1).sum()) for row in got]


# ------------------------------------------------------------------------------------------- filter
def test_filter_traced_condition_is_a_device_compaction(mode):
    m = market.create_economy_model(num_consumers=100_003, num_producers=2_001, config=jx.ModelConfig(seed=8, rng_mode=mode))
    m.run(steps=5)
    c = m.agent_collections["consumers"]
    st = {k: np.array(c.states[k]) for k in c.states}
    cond = lambda s: (s["income"] > 1.05) & (s["savings"] * 2 >= 0.1) & ~(s["utility"] < 0.2)      # noqa: E731
    from jaxabm_b200 import select
    assert len(select.compile_predicate(cond, c._dev.fields[c._tidx])) > 5          # this condition does trace
    f = c.filter(cond)
    mask = (st["income"] > np.float32(1.05)) & (st["savings"] * np.float32(2) >= np.float32(0.1)) & ~(st["utility"] < np.float32(0.2))
    assert 0 < mask.sum() < mask.size and f.num_agents == int(mask.sum())
    for k in st:
        assert np.array_equal(f.states[k], st[k][mask]), k          # same agents, same order: v[mask]
    assert f.model_config is c.model_config
    # the source collection is untouched and still steps
    for k in st:
        assert np.array_equal(c.states[k], st[k]), k
    m.run(steps=1)


def test_filter_vector_and_bool_columns_and_host_mask_fallback(mode):
    m = schelling.create_schelling_model(96, 7000, seed=5, config=jx.ModelConfig(seed=9, rng_mode=mode))
    m.run(steps=3)
    c = m.agent_collections["agents"]
    st = {k: np.array(c.states[k]) for k in c.states}
    f = c.filter(lambda s: (s["position"][:, 0] < 48) & (s["type"] == 1) & s["satisfied"] & (s["moves"] + 1 > 1))
    mask = (st["position"][:, 0] < 48) & (st["type"] == 1) & st["satisfied"] & (st["moves"] + 1 > 1)
    assert 0 < mask.sum() < mask.size
    for k in st:
        assert np.array_equal(f.states[k], st[k][mask]), k
    # NumPy ufuncs on the columns cannot be traced: the mask is evaluated on the host, the compaction still runs
    # on the device and gives the same collection
    g = c.filter(lambda s: np.logical_and(np.asarray(s["position"])[:, 0] < 48, np.asarray(s["type"]) == 1)
                 & np.asarray(s["satisfied"]) & (np.asarray(s["moves"]) + 1 > 1))
    for k in st:
        assert np.array_equal(g.states[k], f.states[k]), k
    with pytest.raises(ValueError):
        c.filter(lambda s: s["type"] > 5)            # nothing matches: AgentCollection(num_agents=0) raises (agent.py:83-84)


def test_filter_matches_oracle_filter(mode):
    """Against the restated reference method (oracle/runtime.py::AgentCollection.filter)."""
    m = market.create_economy_model(num_consumers=3001, num_producers=500, config=jx.ModelConfig(seed=2, rng_mode=mode))
    om = orules.create_economy_model(num_consumers=3001, num_producers=500, config=ort.ModelConfig(seed=2, rng_mode=mode))
    m.run(steps=10), om.run(steps=10)
    f = m.agent_collections["producers"].filter(lambda s: s["profit"] > 0.35)
    of = om.agent_collections["producers"].filter(lambda s: s["profit"] > np.float32(0.35))
    # trajectories agree to 1e-5, so agents whose profit sits within that of the threshold may differ: compare on the
    # device's own columns instead, and the counts loosely
    pr = np.array(m.agent_collections["producers"].states["profit"])
    assert f.num_agents == int((pr > np.float32(0.35)).sum())
    assert abs(f.num_agents - of.num_agents) <= max(3, of.num_agents // 200)
    assert np.array_equal(f.states["capital"], np.array(m.agent_collections["producers"].states["capital"])[pr > np.float32(0.35)])


# ------------------------------------------------------------------------------- state edits between runs
@pytest.mark.parametrize("bands", ["0", "1"])
def test_schelling_uploads_between_runs_with_packed_cell_payload(mode, bands, monkeypatch):
    """The persistent bit-sliced kernel -- and the band kernels with the whole grid as one band (bands = "1") -- keep
    (agent, moves) with the cell and derive 'position' / 'moves' on read.
    Uploading 'moves' between runs refreshes the counts that travel with the cells WITHOUT rebuilding the grid (the
    empty-cell slot order, hence the trajectory, is unchanged); uploading 'position' rebuilds the cell binning -- with the
    other derived column brought up to date first -- exactly like a fresh model created from those columns."""
    monkeypatch.setenv("JXB_GRID_BANDS", bands)

    def build(**kw):
        return schelling.create_schelling_model(1024, 800_000, seed=5, config=jx.ModelConfig(seed=9, rng_mode=mode), **kw)
    a, b = build(), build()
    assert (a._dev.profile()[2] == "grid_shard_sweep_kernel") == (bands == "1")
    a.run(steps=4), b.run(steps=4)
    sa, sb = a.agent_collections["agents"].states, b.agent_collections["agents"].states
    moves_before = np.array(sb["moves"])
    sa["moves"] = np.zeros(800_000, dtype=np.int32)
    ra, rb = a.run(steps=3), b.run(steps=3)
    assert np.array_equal(sa["position"], sb["position"]) and np.array_equal(sa["satisfied"], sb["satisfied"])
    assert np.array_equal(np.array(sa["moves"]) + moves_before, sb["moves"])
    assert [float(v) for v in ra["percent_satisfied"]] == [float(v) for v in rb["percent_satisfied"]]
    # position upload = rebuild of the cell binning from the API columns; 'moves' (derived lazily from the cell
    # payload) must have been brought up to date before the rebuild packs it again
    types, pos, mv = np.array(sb["type"]), np.array(sb["position"]), np.array(sb["moves"])
    sb["position"] = pos
    assert np.array_equal(sb["moves"], mv)
    rb2 = b.run(steps=3)
    assert int(np.array(sb["moves"]).sum() - mv.sum()) == int(rb2["total_moves"][-1]) - int(rb["total_moves"][-1])
    grid = b._dev.download_grid()
    p2 = np.array(sb["position"]).astype(np.int64)
    cells = p2[:, 0] * 1024 + p2[:, 1]
    assert np.unique(cells).size == 800_000 and np.array_equal(grid.reshape(-1)[cells], types)


def test_network_rebuild_after_initialize(mode):
    """Model.add_env_state('network_edges', ...) after initialize() (what Network.add_edge does, agentpy.py:574-582):
    edits are batched on the host, ONE CSR rebuild happens before the next step and the previous CSR goes back to the
    pool; the run then equals the oracle's on the final edge list."""
    n = 30_000
    e1 = synthetic.ring_lattice_edges(n, 2)
    e2 = synthetic.scale_free_edges(n, 3, 11)
    kw = dict(beta=0.25, gamma=0.1, initial_infected=0.02, seed=4)
    m = sir.create_sir_model(n, e1, config=jx.ModelConfig(seed=4, rng_mode=mode), **kw)
    m.initialize()
    launches0 = jx._native.engine().launch_count
    for k in range(5):                                           # five edits, one rebuild
        m.add_env_state("network_edges", e2[: len(e2) // 5 * (k + 1)])
    m.add_env_state("network_edges", e2)
    assert jx._native.engine().launch_count == launches0         # nothing was built yet
    r = m.run(steps=12)
    om = orules.create_sir_model(n, e2, config=ort.ModelConfig(seed=4, rng_mode=mode), **kw)
    orr = om.run(steps=12)
    for k in ("count_S", "count_I", "count_R"):
        assert [int(v) for v in r[k]] == [int(v) for v in orr[k]], k
    assert np.array_equal(m.agent_collections["agents"].states["state"], om.agent_collections["agents"].states["state"])
    # a rebuild in the middle of a run keeps the current epidemic state (it lives in the packed buffers being replaced)
    m.add_env_state("network_edges", np.array(e2))
    r2, or2 = m.run(steps=5), om.run(steps=5)
    assert [int(v) for v in r2["count_I"]] == [int(v) for v in or2["count_I"]]
    assert np.array_equal(m.agent_collections["agents"].states["state"], om.agent_collections["agents"].states["state"])
    # a bad edge list is rejected at the rebuild, with the model still usable afterwards
    bad = np.array(e2)
    bad[7, 1] = n + 5
    m.add_env_state("network_edges", bad)
    with pytest.raises(jx._native.JxbError, match="outside"):
        m.run(steps=1)
    m.add_env_state("network_edges", e2)
    m.run(steps=1)
