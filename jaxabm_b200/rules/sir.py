"""SIR epidemic on a Network (C3).

The reference has no epidemic model; the env layout is
``jaxabm/agentpy.py:557,574-582`` (``network_edges`` int32[E,2], both directions stored for an
undirected network) and the rule is the builder-authored one in DESIGN.md ("SIR rule"),
implemented by ``csrc/sir.cuh`` and restated by ``oracle/rules.py::SIRAgent``.
"""
from __future__ import annotations

import numpy as np

from ..agent import AgentCollection, AgentType
from ..agentpy import Agent, Model as FacadeModel, Network
from ..core import ModelConfig
from ..model import Model
from . import program


class SIRAgent(AgentType):
    """state: ``state`` i32 in {0:S, 1:I, 2:R}."""
    jxb_rule = "sir"

    def __init__(self, beta=0.05, gamma=0.1, initial_infected=0.01):
        self.beta, self.gamma, self.initial_infected = beta, gamma, initial_infected

    def jxb_params(self):
        return [self.beta, self.gamma, self.initial_infected]


class SIRPerson(Agent):
    """Facade agent of the same rule; ``add_agents(n, SIRPerson, beta=..., gamma=..., initial_infected=...)``."""
    jxb_rule = "sir"

    @classmethod
    def jxb_params(cls, p):
        return [p.get("beta", 0.05), p.get("gamma", 0.1), p.get("initial_infected", 0.01)]


@program("sir")
def sir_metrics(env_state, agent_states, params):
    raise RuntimeError("device-resident model function; it is not called on the host")


def create_sir_model(n: int, edges, beta=0.05, gamma=0.1, initial_infected=0.01, seed=42,
                     config: ModelConfig = None) -> Model:
    """``edges``: directed adjacency int32[E,2] exactly as ``Network`` stores it."""
    if config is None:
        config = ModelConfig(seed=seed)
    model = Model(params={"beta": beta, "gamma": gamma}, config=config, metrics_fn=sir_metrics)
    model.add_agent_collection("agents", AgentCollection(SIRAgent(beta, gamma, initial_infected), int(n)))
    model.add_env_state("network_directed", True)
    model.add_env_state("network_edges", np.asarray(edges, dtype=np.int32))
    return model


class SIRModel(FacadeModel):
    """Facade form: parameters ``n_agents``, ``edges`` (undirected pairs int[E,2]) , ``beta``,
    ``gamma``, ``initial_infected``, ``steps``, ``seed``."""
    jxb_program = "sir"

    def setup(self):
        n = self.p.get("n_agents", 100)
        self.agents = self.add_agents(n, SIRPerson, name="agents", beta=self.p.get("beta", 0.05),
                                      gamma=self.p.get("gamma", 0.1),
                                      initial_infected=self.p.get("initial_infected", 0.01))
        self.network = Network(self, directed=False)
        if self.p.get("edges") is not None:
            self.network.add_edges(self.p["edges"])
