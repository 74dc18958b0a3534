"""Sweep model (C5): ``DummyAgent`` + ``create_test_model`` of
``tests/unit/test_analysis.py:22-144`` -> kernel ``rule_growth`` + ``JXB_PROGRAM_GROWTH``."""
from __future__ import annotations

import numpy as np

from ..agent import AgentCollection, AgentType
from ..core import ModelConfig
from ..model import Model
from . import program


class DummyAgent(AgentType):
    """state: value (float32); ``value *= 1 + growth_rate`` each step."""
    jxb_rule = "growth"

    def __init__(self, growth_rate=0.1, initial_value=0.0):
        self.growth_rate = growth_rate
        self.initial_value = initial_value

    def jxb_params(self):
        # the reference multiplies a float32 column by the Python float (1.0 + growth_rate):
        # the factor is formed in double and then rounded once to float32
        return [float(np.float32(1.0 + float(self.growth_rate))), self.initial_value]


GrowthAgent = DummyAgent


@program("growth")
def update_fn(env_state, agent_states, params, key):
    raise RuntimeError("device-resident model function; it is not called on the host")


@program("growth")
def metrics_fn(env_state, agent_states, params):
    raise RuntimeError("device-resident model function; it is not called on the host")


def create_test_model(growth_rate=0.1, adjustment_rate=0.1, initial_value=0.0, num_agents=10, seed=0,
                      params=None, config=None):
    """``test_analysis.py:43-144``."""
    if params is not None:
        growth_rate = params.get("growth_rate", growth_rate)
        adjustment_rate = params.get("adjustment_rate", adjustment_rate)
    if config is None:
        config = ModelConfig(seed=seed)
    model = Model(params={"adjustment_rate": adjustment_rate, "target_price": 1.2}, config=config,
                  update_state_fn=update_fn, metrics_fn=metrics_fn)
    model.add_agent_collection("consumers", AgentCollection(
        DummyAgent(growth_rate=growth_rate, initial_value=initial_value), num_agents))
    model.add_env_state("price_level", 1.0)
    model.add_env_state("interest_rate", 0.05)
    return model
