"""Drop-in for the reference's own example: the user code of examples/basic_example.py --
``jx.Agent.setup/step`` with vector-valued fields, ``jx.Model.setup/step/compute_metrics`` -- runs with only its
two framework imports changed (tests/golden/basic_example_model.py).  Nothing in it is a registered rule: the
facade builds the core model exactly as jaxabm/agentpy.py:1040-1114 does (AgentWrapper, the update_state bridge)
and the rule tracer turns setup / step / update_state / compute_metrics into ONE generated sm_100a kernel.
Checked against the hand-written kernel (rules.random_walk) and the CPU oracle, incl. the facade quirks of
SURVEY.md Appendix B (env overlay; metrics look the walkers up under a fixed name)."""
import importlib.util
import os

import numpy as np
import pytest

from jaxabm_b200.rules import random_walk
from oracle import rules as orules

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
SERIES = ("mean_x", "mean_y", "mean_distance", "max_distance", "num_red", "num_blue", "time")


def _user_module():
    import sys
    sys.path.insert(0, os.path.join(HERE, "examples"))
    import load_example
    load_example.check_against_reference()              # no-op on the GPU box (no reference tree there)
    spec = importlib.util.spec_from_file_location("basic_example_model", load_example.LOCAL)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def series(d, k):
    return np.array([float(v) for v in d[k]], dtype=np.float64)


def test_reference_example_runs_unchanged(mode):
    """The example as shipped: add_agents() names the collection 'randomwalkers', so compute_metrics' lookup of
    'walkers' misses and the distances are the 0.0 defaults (Appendix B) -- identical to the registered model."""
    ex = _user_module()
    p = {"n_agents": 1000, "steps": 100, "seed": 42, "rng_mode": mode}
    model = ex.RandomWalkModel(dict(p))
    res = model.run()
    assert model._jax_model._program == "traced" and "jxc_step_kernel" in model._jax_model._traced_source
    ref = random_walk.RandomWalkModel(dict(p))
    rres = ref.run()
    om = orules.RandomWalkModel({"n_agents": 1000, "steps": 100, "seed": 42}, rng_mode=mode)
    ores = om.run()
    assert list(res._data["step"]) == list(rres._data["step"]) == list(ores["step"])
    for k in SERIES:
        assert np.array_equal(series(res._data, k), series(rres._data, k)), k
        assert np.array_equal(series(res._data, k), series(ores, k)), k
    st, rst = model.walkers.collection.states, ref.walkers.collection.states
    for k in ("position", "velocity", "color", "steps_taken"):
        assert st[k].dtype == rst[k].dtype and st[k].shape == rst[k].shape
        assert np.array_equal(st[k], rst[k]), k                    # elementwise fp32 / int32: bit for bit
        assert np.array_equal(st[k], om.walkers.collection.states[k]), k
    assert not any(k.startswith("agents.") for k in res._data)
    # the host-side part of the example's Model.step(): the core model's env copy counts the steps (model.py:142-144)
    assert model._jax_model.state["env"]["time"] == 100
    # the facade re-runs setup() on every run() (Appendix B): a second run starts over and gives the same series
    res2 = model.run()
    for k in SERIES:
        assert np.array_equal(series(res2._data, k), series(res._data, k)), k


def test_reference_example_with_the_collection_named_walkers(mode):
    """Same user classes, the collection registered as 'walkers': compute_metrics' vector branch --
    sqrt(sum((positions - center) ** 2, axis=1)), mean and max over the agents -- is traced into the kernel's
    reductions."""
    ex = _user_module()

    class Named(ex.RandomWalkModel):
        def setup(self):
            super().setup()
            self._agent_lists.clear()
            self.walkers = self.add_agents(self.p.get("n_agents", 50), ex.RandomWalker, name="walkers")

    p = {"n_agents": 4099, "steps": 120, "seed": 7, "rng_mode": mode}
    model = Named(dict(p))
    res = model.run()
    ref = random_walk.RandomWalkModel(dict(p, name="walkers"))
    rres = ref.run()
    om = orules.RandomWalkModelNamed({"n_agents": 4099, "steps": 120, "seed": 7}, rng_mode=mode)
    ores = om.run()
    for k in ("mean_x", "mean_y", "num_red", "num_blue", "time", "max_distance"):
        assert np.array_equal(series(res._data, k), series(rres._data, k)), k
        assert np.array_equal(series(res._data, k), series(ores, k)), k
    assert series(res._data, "max_distance").max() > 0.6            # the walkers did reach the walls
    # float32 mean: partial sums are folded in a different order than in the hand-written kernel / NumPy
    np.testing.assert_allclose(series(res._data, "mean_distance"), series(rres._data, "mean_distance"), rtol=1e-6)
    np.testing.assert_allclose(series(res._data, "mean_distance"), series(ores, "mean_distance"), rtol=1e-6)
    st, rst = model.walkers.collection.states, ref.walkers.collection.states
    for k in ("position", "velocity", "color", "steps_taken"):
        assert np.array_equal(st[k], rst[k]), k


def test_reference_sensitivity_example_model_with_host_side_step_state(mode):
    """examples/sensitivity/simple_sensitivity_example.py (tests/golden/simple_sensitivity_model.py: one import line
    changed).  Its Model.step() advances host-side state every step (`self.env.add_state('time', self.env.time + 1)`),
    which cannot be fused into the kernel: the traced kernel keeps 'time' as an env READ, and run() performs one device
    step at a time with step() on the host and the refreshed slot before each -- what the reference's un-jitted loop
    does.  Expected values are the closed forms of the example: size_t = prod fl32(1 + g), mean_size stays 1.0 (the
    core model's state never holds 'agents'), efficiency_t = 1 / (t g + 1)."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "examples"))
    import load_example
    load_example.check_against_reference()
    ex = load_example.load(load_example.LOCAL_SENS)
    g, cap, n, T = 0.13, 80.0, 37, 25
    model = ex.SimpleModel({"n_agents": n, "steps": T, "growth_rate": g, "carrying_capacity": cap, "rng_mode": mode})
    res = model.run()
    assert model._jax_model._program == "traced" and model._host_env_keys == {"time"}
    assert [float(v) for v in res._data["final_size"]] == [1.0] * T
    assert [float(v) for v in res._data["resource_usage"]] == [1.0 / cap] * T
    np.testing.assert_allclose([float(v) for v in res._data["efficiency"]], [1.0 / (t * g + 1.0) for t in range(1, T + 1)], rtol=1e-15)
    size = np.float32(1.0)
    for _ in range(T):
        size = np.float32(size * np.float32(1.0 + g))
    st = model.agents.collection.states
    assert np.array_equal(st["size"], np.full(n, size, dtype=np.float32))
    assert np.array_equal(st["growth_rate"], np.full(n, np.float32(g), dtype=np.float32))
    assert model.env.time == T and model._jax_model.state["env"]["time"] == T
    # the example's driver: jx.SensitivityAnalyzer over the model class (host glue around model_class(params).run())
    import jaxabm_b200 as jx
    an = jx.SensitivityAnalyzer(model_class=ex.SimpleModel,
                                parameters=[jx.Parameter("growth_rate", bounds=(0.05, 0.3)),
                                            jx.Parameter("carrying_capacity", bounds=(50.0, 200.0))],
                                n_samples=3, metrics=["final_size", "resource_usage", "efficiency"])
    an.run()
    sens = an.calculate_sensitivity()
    assert set(sens) == {"final_size", "resource_usage", "efficiency"}
