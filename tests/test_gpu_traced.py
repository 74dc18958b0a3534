"""Rule tracer (jaxabm_b200/trace.py): plain user AgentType / update_state_fn / metrics_fn code is traced
into a generated sm_100a step kernel and compared with the SAME user code run eagerly on NumPy columns
by the oracle (oracle/eager.py) on the same seeds."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import jaxabm_b200 as jx  # noqa: E402
from jaxabm_b200 import trace  # noqa: E402
from oracle import rules as orules, runtime as ort  # noqa: E402
from traced_models import build  # noqa: E402

pytestmark = pytest.mark.gpu


def _series(r, k):
    return np.array([float(v) for v in r[k]], dtype=np.float64)


@pytest.mark.parametrize("nc,npr", [(2000, 500), (50_001, 4099)])
def test_traced_market_matches_eager_oracle_and_registered_kernel(mode, nc, npr):
    """The reference's integration-test economy as user code: traced kernel vs eager oracle vs the
    hand-written registered kernel -- float32 trajectories within 1e-5, initial draws bit-exact."""
    m = build.device_market(nc, npr, 3, mode)
    m.initialize()
    o = build.oracle_market(nc, npr, 3, mode)
    o.initialize()
    assert np.array_equal(m.agent_collections["consumers"].states["income"], o.agent_collections["consumers"].states["income"])
    assert np.array_equal(m.agent_collections["producers"].states["capital"], o.agent_collections["producers"].states["capital"])
    r, ro = m.run(steps=25), o.run(steps=25)
    assert list(r.keys()) == list(ro.keys())
    for k in ("gdp", "price_level", "unemployment", "avg_utility", "avg_profit"):
        assert np.allclose(_series(r, k), _series(ro, k), rtol=1e-5, atol=1e-7), k
    from jaxabm_b200.rules import market
    reg = market.create_economy_model(num_consumers=nc, num_producers=npr, config=jx.ModelConfig(seed=3, rng_mode=mode))
    rr = reg.run(steps=25)
    for k in ("gdp", "price_level", "avg_utility", "avg_profit"):
        assert np.allclose(_series(r, k), _series(rr, k), rtol=1e-5, atol=1e-7), k
    assert np.allclose(m.agent_collections["consumers"].states["savings"], o.agent_collections["consumers"].states["savings"],
                       rtol=1e-4, atol=1e-5)
    # env dict is pulled back with the dtypes the reference would hold after a step
    assert isinstance(m._env_state["price_level"], np.float32) and m._env_state["gdp"] == pytest.approx(float(ro["gdp"][-1]), rel=1e-5)


@pytest.mark.parametrize("n", [1000, 100_003])
def test_traced_noisy_traders(mode, n):
    """A model with no registered kernel: per-agent normal + uniform draws in update, bool and int
    state, env-level noise from the update key, sum / mean / max / min reductions of expressions."""
    m, o = build.device_noisy(n, 4, mode), build.oracle_noisy(n, 4, mode)
    r, ro = m.run(steps=8), o.run(steps=8)
    st, ost = m.agent_collections["traders"].states, o.agent_collections["traders"].states
    assert np.array_equal(st["active"], ost["active"]) and np.array_equal(st["trades"], ost["trades"])
    assert np.allclose(st["wealth"], ost["wealth"], rtol=2e-5, atol=1e-5)
    for k in ("n_active", "total_trades", "steps_done"):
        assert [int(v) for v in r[k]] == [int(v) for v in ro[k]], k
    for k in ("mean_wealth", "max_wealth", "min_wealth", "volatility", "participation", "rich_share"):
        assert np.allclose(_series(r, k), _series(ro, k), rtol=2e-5, atol=1e-6), k
    # state persists across run() calls and the second call replays the steady-state variant through graphs
    r2, ro2 = m.run(steps=40), o.run(steps=40)
    assert [int(v) for v in r2["total_trades"]] == [int(v) for v in ro2["total_trades"]]
    assert list(r2["step"]) == list(range(9, 49))


def test_tracer_rejects_what_it_cannot_compile():
    import jaxabm_b200.numpy as jnp
    from jaxabm_b200.agent import AgentCollection, AgentType
    from jaxabm_b200.model import Model

    class Branchy(AgentType):
        def init_state(self, cfg, key):
            return {"x": 1.0}

        def update(self, state, model_state, cfg, key):
            if state["x"] > 0:                      # Python control flow on a traced value
                return {"x": state["x"] + 1.0}
            return {"x": state["x"]}

    m = Model(config=jx.ModelConfig(seed=0))
    m.add_agent_collection("a", AgentCollection(Branchy(), 10))
    with pytest.raises(trace.TraceError):
        m.initialize()


def test_sensitivity_analysis_over_a_traced_factory(mode):
    """SensitivityAnalysis.run with a factory of user-written (traced) models: the float constants live
    in a table, so the sweep compiles ONE kernel and every sample only uploads its values; results
    match the eager oracle sample by sample and move in the direction the reference's integration
    test asserts (tests/integration/test_integration.py:338-348)."""
    import jaxabm_b200.numpy as jnp
    from jaxabm_b200 import jit, random
    from jaxabm_b200.agent import AgentCollection, AgentType
    from jaxabm_b200.analysis import SensitivityAnalysis
    from jaxabm_b200.model import Model
    from oracle import eager
    from traced_models import market as um
    Consumer, Producer, ums, cm = um.make(jnp, random, AgentType)

    def factory(params=None, config=None):
        config.rng_mode = mode
        config.steps = 12
        m = Model(params={"price_adjustment_rate": 0.1}, config=config, update_state_fn=ums, metrics_fn=cm)
        m.add_agent_collection("consumers", AgentCollection(Consumer(propensity_to_consume=params["propensity_to_consume"]), 300))
        m.add_agent_collection("producers", AgentCollection(Producer(productivity=params["productivity"]), 60))
        for k, v in build.MARKET_ENV.items():
            m.add_env_state(k, v)
        return m

    n_before = len(jit._loaded)
    sa = SensitivityAnalysis(factory, {"propensity_to_consume": (0.6, 0.9), "productivity": (0.8, 1.5)}, ["gdp", "avg_utility"],
                             num_samples=6, seed=2)
    res = sa.run(verbose=False)
    assert len(jit._loaded) - n_before <= 1                      # one compiled kernel for the whole sweep
    oC, oP, oums, ocm = um.make(eager.jnp, eager.random, eager.AgentTypeBase)
    for i in range(6):
        p = {k: float(sa.samples[i, j]) for j, k in enumerate(("propensity_to_consume", "productivity"))}
        o = ort.Model(params={"price_adjustment_rate": 0.1}, config=ort.ModelConfig(seed=i + 1000, steps=12, rng_mode=mode),
                      update_state_fn=eager.wrap_model_fn(oums, mode), metrics_fn=eager.wrap_model_fn(ocm, mode, has_key=False))
        o.add_agent_collection("consumers", ort.AgentCollection(eager.wrap_agent_type(oC(propensity_to_consume=p["propensity_to_consume"])), 300))
        o.add_agent_collection("producers", ort.AgentCollection(eager.wrap_agent_type(oP(productivity=p["productivity"])), 60))
        for k, v in build.MARKET_ENV.items():
            o.add_env_state(k, v)
        ro = o.run()
        assert float(res["gdp"][i]) == pytest.approx(float(ro["gdp"][-1]), rel=1e-5)
        assert float(res["avg_utility"][i]) == pytest.approx(float(ro["avg_utility"][-1]), rel=1e-5)
    idx = sa.sobol_indices()
    assert idx["gdp"]["productivity"] > idx["gdp"]["propensity_to_consume"]


def _random_rule(rng, jnp, depth=4):
    """A random expression generator, replayed identically for both backends (same rng stream)."""
    ops = ["add", "sub", "mul", "div", "where", "min", "max", "neg", "abs", "cmp", "const_f", "const_i", "field_x",
           "field_k", "field_b", "sqrt", "log1p", "pow2", "astype_f", "astype_i", "env", "time"]

    def gen(state, env, t, d, want=None):
        choice = ops[rng.randint(len(ops))] if d > 0 else ["const_f", "const_i", "field_x", "field_k", "env"][rng.randint(5)]
        if choice == "const_f":
            return float(np.round(rng.uniform(-2, 2), 3))
        if choice == "const_i":
            return int(rng.randint(-3, 4))
        if choice == "field_x":
            return state["x"]
        if choice == "field_k":
            return state["k"]
        if choice == "field_b":
            return jnp.where(state["b"], gen(state, env, t, d - 1), gen(state, env, t, d - 1))
        if choice == "env":
            return env["scale"]
        if choice == "time":
            return t * 0.125
        a = gen(state, env, t, d - 1)
        if choice == "neg":
            return -a
        if choice == "abs":
            return jnp.abs(a)
        if choice == "sqrt":
            return jnp.sqrt(jnp.abs(a) + 1.0)
        if choice == "log1p":
            return jnp.log1p(jnp.abs(a))
        if choice == "pow2":
            return a ** 2
        if choice == "astype_f":
            return (a * 1).astype(float) if hasattr(a * 1, "astype") else float(a)
        if choice == "astype_i":
            return jnp.clip(a, -1000, 1000).astype(int)
        b = gen(state, env, t, d - 1)
        if choice == "add":
            return a + b
        if choice == "sub":
            return a - b
        if choice == "mul":
            return a * b
        if choice == "div":
            return a / (jnp.abs(b) + 1.5)
        if choice == "min":
            return jnp.minimum(a, b)
        if choice == "max":
            return jnp.maximum(a, b)
        if choice == "cmp":
            return jnp.where(a < b, a, b * 0.5)
        if choice == "where":
            return jnp.where(state["x"] > 0.25, a, b)
        raise AssertionError(choice)
    return gen


@pytest.mark.parametrize("case", range(6))
def test_tracer_dtype_rules_on_random_expressions(mode, case):
    """Random expression trees over float32 / int32 / bool columns, Python constants (weak), an env
    scalar and time_step: the traced kernel and the eager NumPy oracle run the SAME generated user
    code; float columns agree to rounding, integer / boolean columns exactly."""
    import jaxabm_b200.numpy as djnp
    from jaxabm_b200 import random as drandom
    from jaxabm_b200.agent import AgentCollection, AgentType
    from jaxabm_b200.model import Model
    from oracle import eager

    def make(jnp, random, Base):
        class A(Base):
            def init_state(self, cfg, key):
                k1, k2, k3 = random.split(key, 3)
                return {"x": random.uniform(k1, minval=-1.0, maxval=1.0), "k": (random.uniform(k2) * 7.0).astype(int),
                        "b": random.uniform(k3) < 0.5}

            def update(self, state, model_state, cfg, key):
                rng = np.random.RandomState(100 + case)
                gen = _random_rule(rng, jnp)
                env, t = model_state["env"], model_state["time_step"]
                x = jnp.clip(gen(state, env, t, 4) * 1.0, -50.0, 50.0)
                k = jnp.clip(gen(state, env, t, 3), -20, 20)
                b = gen(state, env, t, 3) > gen(state, env, t, 2)
                return {"x": (x * 1.0).astype(float), "k": (k * 1).astype(int), "b": b}

        def upd(env, agent_states, params, key):
            a = agent_states["a"]
            new = dict(env)
            new["scale"] = jnp.clip(jnp.mean(a["x"]) * 0.5 + env["scale"] * 0.5, -3.0, 3.0)
            return new

        def met(env, agent_states, params):
            a = agent_states["a"]
            return {"sx": jnp.sum(a["x"]), "sk": jnp.sum(a["k"]), "nb": jnp.sum(a["b"]), "mx": jnp.max(a["x"]),
                    "scale": env["scale"]}
        return A, upd, met

    n = 4099
    A, upd, met = make(djnp, drandom, AgentType)
    m = Model(config=jx.ModelConfig(seed=21 + case, rng_mode=mode), update_state_fn=upd, metrics_fn=met)
    m.add_agent_collection("a", AgentCollection(A(), n))
    m.add_env_state("scale", 0.75)
    OA, oupd, omet = make(eager.jnp, eager.random, eager.AgentTypeBase)
    o = ort.Model(config=ort.ModelConfig(seed=21 + case, rng_mode=mode), update_state_fn=eager.wrap_model_fn(oupd, mode),
                  metrics_fn=eager.wrap_model_fn(omet, mode, has_key=False))
    o.add_agent_collection("a", ort.AgentCollection(eager.wrap_agent_type(OA()), n))
    o.add_env_state("scale", 0.75)
    r, ro = m.run(steps=5), o.run(steps=5)
    st, ost = m.agent_collections["a"].states, o.agent_collections["a"].states
    assert np.array_equal(st["k"], ost["k"]), np.nonzero(st["k"] != ost["k"])[0][:5]
    assert np.array_equal(st["b"], ost["b"])
    assert np.allclose(st["x"], ost["x"], rtol=1e-4, atol=1e-5, equal_nan=True)
    assert [int(v) for v in r["sk"]] == [int(v) for v in ro["sk"]] and [int(v) for v in r["nb"]] == [int(v) for v in ro["nb"]]
    for k in ("sx", "mx", "scale"):
        assert np.allclose(_series(r, k), _series(ro, k), rtol=1e-4, atol=1e-4), k


def test_traced_ensemble_matches_single_runs(mode):
    """Replica-parallel ensemble kernel of a traced model (generated with the step kernel): its in-kernel
    key schedule (model.py:129-130,156,164,183), per-replica env values and seeds must reproduce the
    individual device runs and the eager oracle."""
    from jaxabm_b200 import ensemble
    from traced_models import build as b
    seeds = [5, 6, 7, 1234]
    vols = [0.02, 0.03, 0.01, 0.05]
    models = []
    for s, v in zip(seeds, vols):
        m = b.device_noisy(3001, s, mode)
        m.add_env_state("volatility", v)
        m.config.steps = 9
        models.append(m)
    assert ensemble.batchable(models)
    last, secs = ensemble.run_last_metrics(models, steps=9)
    assert secs > 0.0
    for i, (s, v) in enumerate(zip(seeds, vols)):
        single = b.device_noisy(3001, s, mode)
        single.add_env_state("volatility", v)
        r = single.run(steps=9)
        o = b.oracle_noisy(3001, s, mode)
        o.add_env_state("volatility", v)
        ro = o.run(steps=9)
        for k in ("n_active", "total_trades", "steps_done"):
            assert int(last[k][i]) == int(r[k][-1]) == int(ro[k][-1]), (k, i)
        for k in ("mean_wealth", "max_wealth", "volatility", "participation", "rich_share"):
            assert float(last[k][i]) == pytest.approx(float(r[k][-1]), rel=2e-5, abs=1e-6), (k, i)
            assert float(last[k][i]) == pytest.approx(float(ro[k][-1]), rel=5e-5, abs=1e-6), (k, i)
