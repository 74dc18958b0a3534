"""CPU tests of the product's host side: the C-ABI library loads and exports every declared
symbol, the host key algebra is bit-compatible, the Python API mirrors the reference's
contract (errors, defaults, bookkeeping) -- no kernel is launched here."""
import os
import re

import numpy as np
import pytest

import jaxabm_b200 as jx
from jaxabm_b200 import _native as nat, dist, ensemble, random as jr
from jaxabm_b200.rules import contract, growth, market, random_walk, schelling, sir
from oracle import jaxlike as jl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "jxb.h")).read()
    declared = set(re.findall(r"\b(jxb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 35
    lib = nat.lib()
    for name in declared:
        assert hasattr(lib, name), f"libjxb.so does not export {name}"
    assert declared == set(nat.SIGNATURES), declared ^ set(nat.SIGNATURES)
    assert lib.jxb_version() == 100


def test_no_device_fails_loudly():
    import ctypes as C
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    h = C.c_void_p()
    rc = nat.lib().jxb_engine_create(0, C.byref(h))
    assert rc == -2 and b"no CPU fallback" in nat.lib().jxb_last_error()
    with pytest.raises(nat.JxbError):
        market.create_economy_model().run(steps=1)


def test_host_prng_matches_oracle_and_kats(mode):
    assert jr.threefry2x32([0, 0], [0, 0]).tolist() == [0x6B200159, 0x99BA4EFE]
    assert jr.threefry2x32([0x13198A2E, 0x03707344], [0x243F6A88, 0x85A308D3]).tolist() == [0xC4923A9C, 0x483DF7A0]
    assert jr.split(jr.PRNGKey(0), 2, 0).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert jr.split(jr.PRNGKey(0), 2, 1).tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]
    # values printed in JAX's documentation (tests/golden/threefry_kat.json)
    assert float(jr.uniform(jr.PRNGKey(0), (), 0.0, 1.0, 0)) == pytest.approx(0.41845703, abs=1e-8)
    assert float(jr.uniform(jr.PRNGKey(0), (), 0.0, 1.0, 1)) == pytest.approx(0.947667, abs=5e-7)
    for seed in (0, 42, 12345):
        k = jr.PRNGKey(seed)
        assert np.array_equal(k, jl.PRNGKey(seed))
        for n in (1, 2, 3, 8, 33):
            assert np.array_equal(jr.split(k, n, mode), jl.split(k, n, mode))
            assert np.array_equal(jr.bits(k, (n,), mode), jl.random_bits(k, (n,), mode))
            assert np.array_equal(jr.uniform(k, (n,), -2.0, 5.0, mode), jl.uniform(k, (n,), -2.0, 5.0, mode))
            assert np.array_equal(jr.randint(k, (n,), 0, 1_000_000, mode), jl.randint(k, (n,), 0, 1_000_000, mode))
        assert np.array_equal(jr.bits(k, (), mode), jl.random_bits(k, (), mode))
        assert np.array_equal(jr.permutation(k, np.arange(37), mode), jl.permutation(k, np.arange(37), mode))


def test_lhs_samples_follow_the_reference_schedule(mode, monkeypatch):
    """analysis.py:67-95 restated with the oracle's jax.random."""
    monkeypatch.setenv("JXB_RNG_MODE", "legacy" if mode == 0 else "partitionable")
    ranges = {"growth_rate": (0.05, 0.2), "adjustment_rate": (0.05, 0.3), "x": (-1.0, 1.0)}
    sa = jx.SensitivityAnalysis(growth.create_test_model, ranges, ["avg_value"], num_samples=17, seed=3)
    key = jl.PRNGKey(3)
    key, sub = jl.split(key, 2, mode)
    pts = np.linspace(0, 1, 18, dtype=np.float32)[:-1]
    pts = (pts + jl.uniform(sub, (17,), mode=mode) / np.float32(17)).astype(np.float32)
    want = np.zeros((17, 3), dtype=np.float32)
    for i in range(3):
        key, sub = jl.split(key, 2, mode)
        want[:, i] = jl.permutation(sub, pts, mode)
    for i, (lo, hi) in enumerate(ranges.values()):
        want[:, i] = want[:, i] * np.float32(hi - lo) + np.float32(lo)
    assert np.array_equal(sa.samples, want)
    for j, (lo, hi) in enumerate(ranges.values()):      # one sample per stratum
        strata = np.floor((np.sort(sa.samples[:, j]) - lo) / (hi - lo) * 17).astype(int)
        assert strata.tolist() == list(range(17))
    with pytest.raises(ValueError):
        sa.sobol_indices()                                   # analysis.py:180-181


def test_api_contract_without_device():
    assert jx.ModelConfig().__dict__ | {"rng_mode": None} == {"seed": 0, "steps": 100, "track_history": True,
                                                              "collect_interval": 1, "rng_mode": None}
    for bad in (0, -3, 2.5, "7"):
        with pytest.raises(ValueError):
            jx.AgentCollection(contract.IncrementAgent(), bad)           # agent.py:83-84
    m = jx.JaxModel()
    with pytest.raises(ValueError):
        m.initialize()                                                   # model.py:125-126
    with pytest.raises(RuntimeError):
        m.step()                                                         # model.py:152-153
    c = jx.AgentCollection(contract.IncrementAgent(), 4)
    with pytest.raises(ValueError):
        c.update({}, jr.PRNGKey(0), jx.ModelConfig())                    # agent.py:150-151
    with pytest.raises(TypeError):
        c.init(jr.PRNGKey(0), {"seed": 0})                               # agent.py:103-104
    assert c.states is None and c.get_states() is None

    class Unknown(jx.AgentType):
        pass

    m = jx.JaxModel()
    m.add_agent_collection("x", jx.AgentCollection(Unknown(), 3))
    with pytest.raises(jx.UnregisteredRuleError):
        m.initialize()
    m = jx.JaxModel(update_state_fn=lambda e, a, p, k: e)
    m.add_agent_collection("consumers", jx.AgentCollection(contract.IncrementAgent(), 3))
    with pytest.raises(jx.UnregisteredRuleError):
        m.initialize()

    class MyModel(jx.Model):
        def step(self):
            pass

    with pytest.raises(ValueError, match="No agent collections"):           # plain Python -> traced path; model.py:125-126
        MyModel({"steps": 1}).run()

    class Mixed(jx.Model):                # a registered agent rule under a plain-Python Model.step cannot be fused
        def setup(self):
            self.add_agents(3, random_walk.RandomWalker)

        def step(self):
            pass

    with pytest.raises(jx.UnregisteredRuleError, match="cannot be mixed"):
        Mixed({"steps": 1}).run()
    with pytest.raises(NotImplementedError):
        jx.CoreModelCalibrator(growth.create_test_model, {"growth_rate": 0.1}, {"avg_value": 1.0}, method="dqn")
    with pytest.raises(ValueError):
        jx.CoreModelCalibrator(growth.create_test_model, {"growth_rate": 0.1}, {"avg_value": 1.0}, method="gradient")
    assert set(jx.__all__) >= {"Agent", "AgentList", "Environment", "Grid", "Network", "Model", "Results",
                               "AgentType", "AgentCollection", "JaxModel", "ModelConfig", "SensitivityAnalysis",
                               "ModelCalibrator", "convert_to_numpy", "format_time", "run_parallel_simulations"}


def test_facade_host_objects():
    class M(jx.Model):
        pass

    m = M({"seed": 5})
    assert m.seed == 5 and m.steps == 100
    net = jx.Network(m, directed=False)
    net.add_edge(0, 1)
    net.add_edge(1, 2)
    net.add_edge(2, 2)
    e = m.env.network_edges
    assert e.dtype == np.int32 and e.tolist() == [[0, 1], [1, 0], [1, 2], [2, 1], [2, 2]]      # agentpy.py:574-582
    assert net.get_neighbors(1).tolist() == [0, 2]
    d = jx.Network(M(), directed=True)
    d.add_edges([[0, 1], [0, 2], [3, 0]])
    assert d.get_neighbors(0).tolist() == [1, 2]
    g = jx.Grid(m, (7, 9), periodic=True)
    assert m.env.grid_shape == (7, 9) and m.env.grid_periodic is True
    al = m.add_agents(6, schelling.SchellingSocialAgent)
    assert al.name == "schellingsocialagents" and len(al) == 6                                 # agentpy.py:960-961
    g.position_agents(al)                                                                      # silently nothing before init
    res = jx.Results({"step": [1, 2], "x": [0.5, 0.25], "agents.a.v": [np.zeros(3), np.ones(3)]})
    assert "x" in res and res["x"][-1] == 0.25 and len(res.variables.a.v) == 2
    assert jx.format_time(12.345) == "12.35s" and jx.format_time(125) == "2m 5.00s" and jx.format_time(3723) == "1h 2m 3.00s"
    p = jx.Parameter("a", bounds=(0.0, 1.0))
    s = jx.Sample({"a": p, "b": 3}, n=4, seed=1)
    assert len(s) == 4 and all(0 <= x["a"] <= 1 and x["b"] == 3 for x in s)
    # the reference's own call forms (agentpy.py:1230,1247-1267,1294-1311): a list of Parameters, n_samples, indexing
    p2 = jx.Parameter("growth", (0.01, 0.1))
    s2 = jx.Sample([p, p2], n_samples=5)
    assert len(s2) == 5 and s2.n_samples == 5 and set(s2[3]) == {"a", "growth"} and 0.01 <= s2[3]["growth"] <= 0.1
    assert len(s2._samples["a"]) == 5 and s2.parameters[1] is p2
    an2 = jx.SensitivityAnalyzer(random_walk.RandomWalkModel, [p, p2], n_samples=3, metrics=["mean_distance"])
    assert an2.ranges == {"a": (0.0, 1.0), "growth": (0.01, 0.1)} and len(an2.sample) == 3 and an2.fixed == {}
    cal = jx.ModelCalibrator(random_walk.RandomWalkModel, [p2], {"mean_distance": 0.1})
    assert cal._pd == {"growth": p2} and cal.fixed == {}
    with pytest.raises(AttributeError):
        an = jx.SensitivityAnalyzer(random_walk.RandomWalkModel, {"n_agents": p}, n_samples=2)
        an._sa = object.__new__(jx.SensitivityAnalysis)
        an._sa.results = {}
        an.calculate_sensitivity("morris")                                                     # agentpy.py:1363-1364


def test_ensemble_plan_is_pure_host_logic():
    models = [growth.create_test_model(initial_value=1.0, num_agents=50,
                                       params={"growth_rate": 0.05 + 0.01 * i, "adjustment_rate": 0.1 * (i % 3 + 1)},
                                       config=jx.ModelConfig(seed=1000 + i)) for i in range(6)]
    assert ensemble.batchable(models)
    desc, slots, params, seeds, env0 = ensemble.plan(models)
    assert desc.program == nat.PROGRAM["growth"] and desc.n_types == 1 and desc.types[0].n_agents == 50
    assert slots == [0, 100] and params.shape == (6, 2) and seeds.tolist() == list(range(1000, 1006))
    assert params[:, 1].tolist() == [float(np.float32(1.0 + 0.05 + 0.01 * i)) for i in range(6)]
    assert env0[:2].tolist() == [1.0, 0.05]
    mixed = models[:2] + [market.create_economy_model()]
    assert not ensemble.batchable(mixed)
    assert not ensemble.batchable([object()])
    mk = [market.create_economy_model(params={"productivity": 1.0 + 0.1 * i}, config=jx.ModelConfig(seed=i)) for i in range(3)]
    _, slots, params, _, _ = ensemble.plan(mk)
    assert slots == [100 + 16 + 1] and np.allclose(params[:, 0], [1.0, 1.1, 1.2])


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 8192, 8191):
        for w in (1, 2, 3, 8):
            b = [dist.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    assert dist.rank_world() == (0, 1) and dist.max_over_ranks(3.5) == 3.5
    a = np.arange(6.0).reshape(3, 2)
    assert dist.gather_rows(a, 3) is a


def test_synthetic_generators():
    from jaxabm_b200 import synthetic
    e = synthetic.scale_free_edges(5000, 4, 1)
    assert e.dtype == np.int32 and e.shape[1] == 2 and e.min() >= 0 and e.max() < 5000
    assert np.array_equal(np.sort(e[: len(e) // 2], axis=0)[:, ::1], np.sort(e[len(e) // 2:][:, ::-1], axis=0))
    deg = np.bincount(e[:, 0], minlength=5000)
    assert deg.max() > 20 * np.median(deg)                                                     # heavy tail
    assert np.array_equal(e, synthetic.scale_free_edges(5000, 4, 1))
    t, p = schelling.initial_layout(32, 700, 0.5, 3)
    assert len({(a, b) for a, b in p.tolist()}) == 700 and t.sum() == 350


def test_saltelli_estimators_on_an_analytic_model():
    """run_saltelli with a duck-typed host model (the ensemble factory protocol of analysis.py:128-145):
    y = a + 2 b has S1 = ST = (1/5, 4/5) for independent uniform a, b on [0, 1]."""
    from jaxabm_b200.analysis import SensitivityAnalysis

    class Lin:
        def __init__(self, params, config):
            self.p = params

        def run(self, steps=None):
            return {"y": [self.p["a"] + 2.0 * self.p["b"]]}

    sa = SensitivityAnalysis(lambda params, config: Lin(params, config), {"a": (0.0, 1.0), "b": (0.0, 1.0)}, ["y"],
                             num_samples=16, seed=3)
    idx = sa.run_saltelli(num_base=512)
    assert sa.saltelli_design.shape == (512 * 4, 2)
    assert idx["y"]["S1"]["a"] == pytest.approx(0.2, abs=0.04) and idx["y"]["S1"]["b"] == pytest.approx(0.8, abs=0.06)
    assert idx["y"]["ST"]["a"] == pytest.approx(0.2, abs=0.03) and idx["y"]["ST"]["b"] == pytest.approx(0.8, abs=0.05)


def test_rule_tracer_generates_and_compiles_a_kernel_without_a_gpu():
    """Plain user code -> expression graph -> CUDA source -> sm_100a library (nvcc cross-compiles here);
    dtype rules of the trace follow JAX's weak typing."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from jaxabm_b200 import jit, trace
    from traced_models import build
    m = build.device_noisy(1000, 1, 1)
    variants = jit.trace_variants(m)
    assert [v.env_dtypes for v in variants] == [["wf64", "wf64", "wi32"], ["f32", "f32", "wi32"]]
    assert variants[-1].types[0]["fields"] == [("wealth", "f32", 1), ("active", "bool", 1), ("trades", "i32", 1)]
    src, meta = trace.generate_source(variants)
    assert "normal_scalar<MODE>" in src and "jxc_step_kernel" in src and meta["n_variants"] == 2
    lib = jit.compile_source(src)
    assert lib.jxc_n_acc() == meta["n_acc"] >= 6 and lib.jxc_n_variants() == 2
    # weak-type rules
    x = trace.Tr("field", (), trace.I32, "a", (0, 0))
    assert (x * 0.5).dtype == trace.F32 and (x * 2).dtype == trace.I32 and (x / 2).dtype == trace.F32
    assert (x > 1).dtype == trace.BOOL and trace.where(x > 1, 1.0, 0.0).dtype == trace.F32
    assert (trace._const(2.0) * 3).attr == 6.0                      # scalar (op) scalar folds in Python doubles
    with pytest.raises(trace.TraceError):
        bool(x > 1)


def test_network_cuts_balance_the_adjacency():
    """Node-range boundaries of a sharded Network (sharding.network_cuts): multiples of 32, every rank
    non-empty, adjacency (+ a per-row cost) balanced; falls back to an even split without edges."""
    from jaxabm_b200 import sharding, synthetic
    n = 50_000
    e = synthetic.scale_free_edges(n, 5, 3)
    deg = np.bincount(e[:, 0], minlength=n)
    for world in (2, 3, 8):
        c = sharding.network_cuts(n, e, world, balance="entries")
        assert sharding.network_cuts(n, e, world)[1] == (((n + 31) // 32 + world - 1) // world) * 32       # default: even split
        assert c[0] == 0 and c[-1] == n and len(c) == world + 1
        assert all(b > a for a, b in zip(c, c[1:])) and all(x % 32 == 0 for x in c[:-1])
        cost = [int(deg[a:b].sum()) + 4 * (b - a) for a, b in zip(c, c[1:])]
        assert max(cost) < 1.1 * (sum(cost) / world), (world, cost)
        even = [int(deg[a:b].sum()) for a, b in zip(range(0, n, n // world), range(n // world, n + 1, n // world))]
        assert max(even) > 1.25 * (sum(even) / world)          # what the balancing is for: hubs sit at low indices
    assert sharding.network_cuts(100, None, 2) == [0, 64, 100]
    assert sharding.network_cuts(70, e[:0], 3) == [0, 32, 64, 70]


def test_run_parallel_simulations_follows_the_reference_schedule(capsys):
    """jaxabm/utils.py:128-175: seeds seed_offset + i * num_runs + j, 'params' / 'seed' added to every results
    dict, a failing run is reported and skipped."""
    class Fake:
        def __init__(self, params, config):
            self.p, self.c = params, config

        def run(self, steps=None):
            if self.p.get("boom"):
                raise RuntimeError("x")
            return {"step": [steps or 1], "v": [self.p["a"] * 10 + self.c.seed]}

    r = jx.run_parallel_simulations(lambda params, config: Fake(params, config),
                                    [{"a": 1}, {"a": 2, "boom": 1}, {"a": 3}], num_runs=2, seed_offset=100)
    assert [(x["seed"], x["v"], x["params"]["a"]) for x in r] == [(100, [110], 1), (101, [111], 1), (104, [134], 3), (105, [135], 3)]
    out = capsys.readouterr().out
    assert "Running simulation 2/3, run 2/2, seed=103" in out and "Error in simulation 2/3, run 1/2: x" in out
    r = jx.run_parallel_simulations(lambda params, config: Fake(params, config), [{"a": 1}, {"a": 3}], seeds=[7, 9], steps=4)
    assert [(x["seed"], x["step"]) for x in r] == [(7, [4]), (9, [4])]


def test_header_is_plain_c():
    """include/jxb.h is the drop-in boundary: it must compile as C99 on its own, without warnings."""
    import shutil
    import subprocess
    import tempfile
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    with tempfile.NamedTemporaryFile("w", suffix=".c", delete=False) as f:
        f.write('#include "jxb.h"\nint main(void) { return JXB_VERSION == 100 ? 0 : 1; }\n')
    out = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                          "-I", os.path.join(ROOT, "include"), f.name], capture_output=True, text=True)
    os.unlink(f.name)
    assert out.returncode == 0, out.stderr


def test_c_program_links_against_the_library(tmp_path):
    """tests/abi/abi_smoke.c: the C ABI used from C -- host key algebra works, device calls fail loudly without a GPU."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    libdir = os.path.join(ROOT, "jaxabm_b200", "csrc")
    exe = str(tmp_path / "abi_smoke")
    build = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                            os.path.join(ROOT, "tests", "abi", "abi_smoke.c"), "-o", exe, "-L", libdir, "-ljxb",
                            f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and "abi ok" in run.stdout, (run.returncode, run.stdout, run.stderr)


def test_filter_condition_compiles_to_a_postfix_program():
    """jaxabm_b200/select.py (AgentCollection.filter, jaxabm/agent.py:213-243): the traced condition as a postfix
    program, checked by interpreting it here agent by agent against the NumPy mask of the same condition."""
    from jaxabm_b200 import select
    from jaxabm_b200.trace import TraceError
    rng = np.random.RandomState(0)
    n = 500
    cols = {"wealth": rng.uniform(0, 20, n).astype(np.float32), "kind": rng.randint(0, 3, n).astype(np.int32),
            "pos": rng.randint(0, 50, (n, 2)).astype(np.int32), "ok": rng.rand(n) < 0.5}
    fields = [("wealth", np.float32, 1), ("kind", np.int32, 1), ("pos", np.int32, 2), ("ok", np.bool_, 1)]
    names = {v: k for k, v in select.OP.items()}

    def run(prog, i):
        st = []
        f32, i32 = np.float32, np.int32
        for op, a, b, f in prog:
            o = names[op]
            if o.startswith("LOAD"):
                col = cols[fields[a][0]]
                st.append(col[i] if col.ndim == 1 else col[i, b])
            elif o == "CONST_F32":
                st.append(f32(f))
            elif o == "CONST_I32":
                st.append(i32(a))
            elif o in ("NOT",):
                st.append(not st.pop())
            elif o in ("I2F", "B2F"):
                st.append(f32(st.pop()))
            elif o in ("F2I", "B2I"):
                st.append(i32(st.pop()))
            elif o in ("I2B", "F2B"):
                st.append(bool(st.pop()))
            elif o[:3] in ("NEG", "ABS"):
                v = st.pop()
                st.append(-v if o[:3] == "NEG" else abs(v))
            elif o == "SELECT":
                y, x, c = st.pop(), st.pop(), st.pop()
                st.append(x if c else y)
            else:
                y, x = st.pop(), st.pop()
                k = o.split("_")[0]
                st.append({"ADD": lambda: x + y, "SUB": lambda: x - y, "MUL": lambda: x * y, "DIV": lambda: x / y,
                           "MIN": lambda: min(x, y), "MAX": lambda: max(x, y), "LT": lambda: x < y, "LE": lambda: x <= y,
                           "GT": lambda: x > y, "GE": lambda: x >= y, "EQ": lambda: x == y, "NE": lambda: x != y,
                           "AND": lambda: bool(x) and bool(y), "OR": lambda: bool(x) or bool(y),
                           "XOR": lambda: bool(x) != bool(y)}[k]())
        assert len(st) == 1
        return bool(st[0])

    conds = [lambda s: s["wealth"] > 10,
             lambda s: (s["wealth"] * 0.5 + 1 >= 4.25) & (s["kind"] == 1),
             lambda s: (s["pos"][:, 0] + s["pos"][:, 1] < 40) | ~s["ok"],
             lambda s: jx.numpy.where(s["ok"], s["wealth"], -s["wealth"]) > 3,
             lambda s: (s["kind"] * 2 - 1 > 0) ^ (abs(s["wealth"] - 10) <= 2.5)]
    for c in conds:
        prog = select.compile_predicate(c, fields)
        want = np.asarray(c({k: (v if k != "ok" else v) for k, v in cols.items()})) if c is not conds[3] else \
            np.where(cols["ok"], cols["wealth"], -cols["wealth"]) > 3
        got = np.array([run(prog, i) for i in range(n)])
        assert np.array_equal(got, want)
    with pytest.raises(TraceError):
        select.compile_predicate(lambda s: np.logical_and(s["wealth"] > 1, s["ok"]), fields)     # a NumPy ufunc on symbols
    with pytest.raises(TraceError):
        select.compile_predicate(lambda s: s["wealth"] + 1, fields)                               # not boolean


def test_feistel_inverse_is_the_inverse_permutation():
    """The Schelling kernels walk the unsatisfied agents in cell order and ask which mover index each one is:
    jxb_prng_feistel(..., inverse=1) must undo the forward matching permutation (cycle walking included), and the
    forward direction must agree with the oracle's restatement."""
    import ctypes as C
    from jaxabm_b200 import _native as nat
    from oracle import jaxlike as jl
    lib = nat.lib()
    rng = np.random.RandomState(1)
    for n in (1, 2, 3, 5, 64, 1000, 4097, 65536, 3_777_216):
        rk = np.ascontiguousarray(rng.randint(0, 2**32, 4, dtype=np.uint64).astype(np.uint32))
        idx = np.arange(n) if n <= 5000 else rng.randint(0, n, 3000)
        fwd = np.empty(len(idx), dtype=np.uint32)
        out = C.c_uint32()
        for i, k in enumerate(idx):
            nat.check(lib.jxb_prng_feistel(n, nat.ptr(rk), int(k), 0, C.byref(out)))
            fwd[i] = out.value
            nat.check(lib.jxb_prng_feistel(n, nat.ptr(rk), int(fwd[i]), 1, C.byref(out)))
            assert out.value == k
        assert np.array_equal(fwd, jl.feistel_permute(np.asarray(idx, dtype=np.uint32), n, rk))
        if n <= 5000:
            assert np.array_equal(np.sort(fwd), np.arange(n))


def test_facade_example_traces_and_compiles_without_a_gpu():
    """The reference's basic example (tests/golden/basic_example_model.py = its two classes with the framework
    imports changed) through the facade's plain-Python path: AgentWrapper + the update_state bridge are traced,
    vector-valued fields become width-2 columns, the kernel source compiles for sm_100a."""
    import importlib.util
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "examples"))
    import load_example
    assert load_example.check_against_reference() in (True, False)
    spec = importlib.util.spec_from_file_location("basic_example_model", load_example.LOCAL)
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    from jaxabm_b200 import jit, trace as T
    from jaxabm_b200.model import Model as CoreModel
    m = ex.RandomWalkModel({"n_agents": 1000, "steps": 100, "seed": 42})
    m.setup()
    assert m._program() is None                       # nothing registered: the traced path
    jm = CoreModel(params=m.p, config=jx.ModelConfig(steps=100, seed=42), update_state_fn=m.update_state,
                   metrics_fn=m.compute_metrics)
    m._jax_model = jm
    for name, al in m._agent_lists.items():
        jm.add_agent_collection(name, al.collection)
    for name, value in m.env.state.items():
        jm.add_env_state(name, value)
    variants = jit.trace_variants(jm)
    tm = variants[-1]
    assert tm.types[0]["name"] == "randomwalkers"
    assert tm.types[0]["fields"] == [("position", "f32", 2), ("velocity", "f32", 2), ("color", "i32", 1), ("steps_taken", "i32", 1)]
    assert [k for k, _ in tm.metrics] == ["mean_x", "mean_y", "mean_distance", "max_distance", "num_red", "num_blue", "time"]
    assert m._step_has_host_effects and m._jax_model is jm        # step()'s add_env_state hit the probe, not the model
    assert jm._env_state["time"] == 0
    src, meta = T.generate_source(variants)
    assert "float4 F0_0" in src and "float4 F0_1" in src          # f32[N,2]: two 16-byte vectors per four agents
    lib = jit.compile_source(src)
    assert lib.jxc_n_variants() == len(variants)


def test_facade_step_with_host_side_state_is_not_frozen_into_the_kernel():
    """A Model.step() that advances Environment.state on the host every step (the reference's sensitivity example)
    must not be traced as a constant: the entry stays an env read and the model is flagged for per-step host calls."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "examples"))
    import load_example
    ex = load_example.load(load_example.LOCAL_SENS)
    from jaxabm_b200 import jit
    from jaxabm_b200.model import Model as CoreModel
    m = ex.SimpleModel({"n_agents": 5, "steps": 3, "growth_rate": 0.2})
    m.setup()
    m._host_env_keys, m._step_has_host_effects = set(), False
    jm = CoreModel(params=m.p, config=jx.ModelConfig(steps=3), update_state_fn=m.update_state, metrics_fn=m.compute_metrics)
    m._jax_model = jm
    for name, al in m._agent_lists.items():
        jm.add_agent_collection(name, al.collection)
    for name, value in m.env.state.items():
        jm.add_env_state(name, value)
    variants = jit.trace_variants(jm)
    assert m._host_env_keys == {"time"} and m.env.state["time"] == 0           # the trace left the host state alone
    assert "time" not in variants[-1].env_out and "growth_rate" in variants[-1].env_out
    assert dict(variants[-1].metrics)["efficiency"].op != "const"               # 1 / (time * g + 1) reads the env slot


def test_handle_sizes_match_the_header():
    """The host shim sizes the buffers it all-gathers from constants that must follow include/jxb.h."""
    import re
    from jaxabm_b200 import _native as nat
    header = open(os.path.join(ROOT, "include", "jxb.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(JXB_\w+_HANDLE_BYTES)\s+(\d+)", header)}
    assert defs == {"JXB_IPC_HANDLE_BYTES": nat.IPC_HANDLE_BYTES, "JXB_GRID_HANDLE_BYTES": nat.GRID_HANDLE_BYTES}
