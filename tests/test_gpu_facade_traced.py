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
