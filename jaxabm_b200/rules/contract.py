"""The small models the reference's unit tests are written against.

* ``IncrementAgent``  = ``DummyAgent`` of ``tests/unit/test_model.py:43-48``
* ``update_state_fn`` / ``metrics_fn`` = ``tests/unit/test_model.py:20-40``
* ``WealthAgent``     = ``TestAgent`` of ``tests/unit/test_agent.py:44-100``
"""
from __future__ import annotations

from ..agent import AgentType
from . import program


class IncrementAgent(AgentType):
    """state: value ~ U(0,10); ``value += env['increment']`` (default 1.0)."""
    jxb_rule = "increment"


class WealthAgent(AgentType):
    """state: wealth ~ U(0,100), productivity ~ U(0.5,1.5); ``wealth += productivity * wage_rate``
    with ``wage_rate = model_state.get('wage_rate', 1.0)`` (test_agent.py:75-78)."""
    jxb_rule = "wealth"

    def jxb_params(self):
        return [1.0]

    def jxb_bind_model_state(self, dev, tidx, model_state):
        dev.set_type_param(tidx, 0, float(model_state.get("wage_rate", 1.0)))


@program("counter")
def update_state_fn(env_state, agent_states, params, key):
    raise RuntimeError("device-resident model function; it is not called on the host")


@program("counter")
def metrics_fn(env_state, agent_states, params):
    raise RuntimeError("device-resident model function; it is not called on the host")
