"""Consumer / producer market (C4-A): host handles for the kernels ``rule_consumer`` /
``rule_producer`` and the ``JXB_PROGRAM_MARKET`` tail (``csrc/rules.cuh``).

Mirrors ``tests/integration/test_integration.py``: ``Consumer`` (:20-67), ``Producer``
(:70-121), ``update_model_state`` (:125-160), ``compute_metrics`` (:163-183),
``create_economy_model`` (:187-283).
"""
from __future__ import annotations

from ..agent import AgentCollection, AgentType
from ..core import ModelConfig
from ..model import Model
from . import program


class Consumer(AgentType):
    """state: savings, consumption, utility, income (float32)."""
    jxb_rule = "consumer"

    def __init__(self, base_income=1.0, propensity_to_consume=0.8):
        self.base_income = base_income
        self.propensity_to_consume = propensity_to_consume

    def jxb_params(self):
        return [self.base_income, self.propensity_to_consume]


class Producer(AgentType):
    """state: capital, production, profit (float32)."""
    jxb_rule = "producer"

    def __init__(self, initial_capital=10.0, productivity=1.0, reinvestment_rate=0.3):
        self.initial_capital = initial_capital
        self.productivity = productivity
        self.reinvestment_rate = reinvestment_rate

    def jxb_params(self):
        return [self.initial_capital, self.productivity, self.reinvestment_rate]


@program("market")
def update_model_state(env_state, agent_states, params, key):
    """Runs on the device as the tail of the step kernel (``program_tail``, MARKET)."""
    raise RuntimeError("device-resident model function; it is not called on the host")


@program("market")
def compute_metrics(env_state, agent_states, params):
    raise RuntimeError("device-resident model function; it is not called on the host")


INITIAL_ENV = {"price_level": 1.0, "gdp": 0.0, "unemployment": 0.0,
               "total_consumption": 0.0, "total_production": 0.0}


def create_economy_model(num_consumers=20, num_producers=5, base_income=1.0, propensity_to_consume=0.8,
                         initial_capital=10.0, productivity=1.0, reinvestment_rate=0.3,
                         price_adjustment_rate=0.1, target_price=1.0, seed=0, params=None, config=None):
    """``test_integration.py:187-283`` (a collection with a zero count is simply omitted)."""
    if params is not None:
        propensity_to_consume = params.get("propensity_to_consume", propensity_to_consume)
        productivity = params.get("productivity", productivity)
        price_adjustment_rate = params.get("price_adjustment_rate", price_adjustment_rate)
    if config is None:
        config = ModelConfig(seed=seed)
    model = Model(params={"price_adjustment_rate": price_adjustment_rate, "target_price": target_price},
                  config=config, update_state_fn=update_model_state, metrics_fn=compute_metrics)
    if num_consumers:
        model.add_agent_collection("consumers", AgentCollection(
            Consumer(base_income=base_income, propensity_to_consume=propensity_to_consume), num_consumers))
    if num_producers:
        model.add_agent_collection("producers", AgentCollection(
            Producer(initial_capital=initial_capital, productivity=productivity,
                     reinvestment_rate=reinvestment_rate), num_producers))
    for name, value in INITIAL_ENV.items():
        model.add_env_state(name, value)
    return model
