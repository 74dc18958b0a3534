"""Schelling segregation on a Grid (C2).

State and env layout follow ``examples/models/schelling_model.py:26-31,119-139``; the
reference's rule body is a placeholder, so the rule is the builder-authored
one in DESIGN.md ("Schelling rule"), implemented by ``csrc/schelling.cuh`` and restated by
``oracle/rules.py::SchellingAgent`` for the parity tests.
"""
from __future__ import annotations

import numpy as np

from ..agent import AgentCollection, AgentType
from ..agentpy import Agent, Grid, Model as FacadeModel
from ..core import ModelConfig
from ..model import Model
from . import program


class SchellingSocialAgent(Agent):
    """state: type i32, position i32[2], satisfied bool, moves i32 (schelling_model.py:26-31)."""
    jxb_rule = "schelling"

    def setup(self):
        return {"type": 0, "position": np.zeros(2, dtype=np.int32), "satisfied": False, "moves": 0}


class SchellingAgentType(AgentType):
    """Core-protocol handle of the same rule."""
    jxb_rule = "schelling"

    def __init__(self, similarity_threshold=0.5):
        self.similarity_threshold = similarity_threshold


@program("schelling")
def schelling_update_state(env_state, agent_states, params, key):
    raise RuntimeError("device-resident model function; it is not called on the host")


@program("schelling")
def schelling_metrics(env_state, agent_states, params):
    raise RuntimeError("device-resident model function; it is not called on the host")


def initial_layout(grid_size: int, n_agents: int, ratio: float, seed: int):
    """Unique random cells for the agents and types by ``ratio``
    (``schelling_model.py:86-95,141-170``): the first N entries of a seeded permutation of the
    cell ids -- the 1-D form of the example's shuffle of the (x, y) table."""
    rng = np.random.RandomState(seed)
    cells = rng.permutation(grid_size * grid_size)[:n_agents]
    pos = np.stack([cells // grid_size, cells % grid_size], axis=1).astype(np.int32)
    n0 = int(n_agents * ratio)
    types = np.concatenate([np.zeros(n0, dtype=np.int32), np.ones(n_agents - n0, dtype=np.int32)])
    return types, pos


class SchellingModel(FacadeModel):
    """``schelling_model.py:72-196`` with a working rule.  Parameters: ``grid_size``,
    ``n_agents``, ``ratio``, ``similarity_threshold``, ``periodic``, ``steps``, ``seed``."""
    jxb_program = "schelling"

    def setup(self):
        g = self.p.get("grid_size", 20)
        self.grid = Grid(self, (g, g), periodic=self.p.get("periodic", False))
        n = self.p.get("n_agents", 300)
        # named 'agents': the key compute_metrics reads (schelling_model.py:175)
        self.agents = self.add_agents(n, SchellingSocialAgent, name="agents")
        self._types, self._positions = initial_layout(g, n, self.p.get("ratio", 0.5), self.p.get("seed", 42))
        self.env.add_state("segregation_index", 0.0)
        self.env.add_state("percent_satisfied", 0.0)
        self.env.add_state("total_moves", 0)

    def after_initialize(self):
        st = self.agents.collection.states
        st["type"] = self._types
        self.grid.position_agents(self.agents, self._positions)

    @property
    def grid_state(self) -> np.ndarray:
        """env['grid'] as the example lays it out: int32[W,H], -1 empty else the agent type."""
        return self._jax_model._dev.download_grid()


def create_schelling_model(grid_size=20, n_agents=300, ratio=0.5, similarity_threshold=0.5, periodic=False,
                           seed=42, config: ModelConfig = None, types=None, positions=None,
                           shard=False) -> Model:
    """Core-protocol construction; returns an *initialized* model with the layout uploaded.

    ``shard=True`` splits the ONE grid into row bands over the ranks of the ``torch.distributed``
    world (``csrc/grid_shard.cuh``); every rank passes the same full layout."""
    if config is None:
        config = ModelConfig(seed=seed)
    if types is None or positions is None:
        types, positions = initial_layout(grid_size, n_agents, ratio, seed)
    coll = AgentCollection(SchellingAgentType(similarity_threshold), int(n_agents))
    model = Model(params={"similarity_threshold": similarity_threshold}, config=config,
                  update_state_fn=schelling_update_state, metrics_fn=schelling_metrics)
    model.add_agent_collection("agents", coll)
    model.add_env_state("grid_shape", (grid_size, grid_size))
    model.add_env_state("grid_periodic", periodic)
    model.add_env_state("segregation_index", 0.0)
    model.add_env_state("percent_satisfied", 0.0)
    model.add_env_state("total_moves", 0)
    if shard:
        from .. import sharding
        sharding.shard_model(model)
    model.initialize()
    coll.states["type"] = types
    coll.states["position"] = positions
    return model
