"""Random walk (C1): ``examples/basic_example.py`` -> kernel ``rule_walker`` +
``JXB_PROGRAM_RANDOM_WALK``.

``RandomWalkModel`` reproduces the example including its observable quirks (SURVEY.md
section 3.1 / Appendix B): the facade's env overlay freezes ``time``/``mean_x``/... at their
``setup()`` values, and ``compute_metrics`` looks for a collection called ``'walkers'`` while
``add_agents`` auto-names it ``'randomwalkers'`` -- so the distance metrics are 0.0 unless
the collection is added with ``name='walkers'`` (``RandomWalkModel(..., {'name': 'walkers'})``).
"""
from __future__ import annotations

import numpy as np

from ..agent import AgentCollection, AgentType
from ..agentpy import Agent, Model as FacadeModel
from ..core import ModelConfig
from ..model import Model
from . import program


class RandomWalker(Agent):
    """``basic_example.py:20-69``: bounce between env['bounds'], flip colour on a bounce."""
    jxb_rule = "random_walker"

    def setup(self):
        return {"position": np.array([0.5, 0.5], dtype=np.float32),
                "velocity": np.array([0.01, 0.01], dtype=np.float32),
                "color": 0, "steps_taken": 0}


class RandomWalkModel(FacadeModel):
    """``basic_example.py:72-182``."""
    jxb_program = "random_walk"
    jxb_metrics_collection = "walkers"

    def setup(self):
        n_agents = self.p.get("n_agents", 50)
        self.walkers = self.add_agents(n_agents, RandomWalker, name=self.p.get("name"))
        self.env.add_state("bounds", np.array([0.0, 1.0], dtype=np.float32))
        self.env.add_state("time", 0)
        self.env.add_state("mean_x", 0.5)
        self.env.add_state("mean_y", 0.5)
        self.env.add_state("num_red", n_agents)
        self.env.add_state("num_blue", 0)


class ScaledRandomWalker(AgentType):
    """Core-protocol walker with a keyed initial state (SURVEY.md 8(d), C1 scaled variant):
    position ~ U(0,1)^2 and velocity ~ U(-0.01,0.01)^2 from the agent's own key."""
    jxb_rule = "scaled_walker"


@program("random_walk")
def walk_update_state(env_state, agent_states, params, key):
    raise RuntimeError("device-resident model function; it is not called on the host")


@program("random_walk")
def walk_metrics(env_state, agent_states, params):
    raise RuntimeError("device-resident model function; it is not called on the host")


def create_scaled_walk_model(n_agents: int, seed: int = 42, config: ModelConfig = None) -> Model:
    """The N=2^26 roofline workload: ``ScaledRandomWalker`` x n under the random-walk program
    (collection named 'walkers' so the fused distance reductions are live)."""
    if config is None:
        config = ModelConfig(seed=seed)
    m = Model(params={}, config=config, update_state_fn=walk_update_state, metrics_fn=walk_metrics)
    m.add_agent_collection("walkers", AgentCollection(ScaledRandomWalker(), n_agents))
    m.add_env_state("bounds", np.array([0.0, 1.0], dtype=np.float32))
    for k, v in (("time", 0), ("mean_x", 0.5), ("mean_y", 0.5), ("num_red", n_agents), ("num_blue", 0)):
        m.add_env_state(k, v)
    return m
