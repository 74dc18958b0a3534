"""Builders shared by the traced-model tests: the same user code for the CUDA tracer and for the
NumPy batch backend of the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traced_models import market as user_models  # noqa: E402

MARKET_ENV = {"price_level": 1.0, "gdp": 0.0, "unemployment": 0.0, "total_consumption": 0.0, "total_production": 0.0}
NOISY_ENV = {"volatility": 0.02, "participation": 0.0, "steps_done": 0}


def device_market(nc, npr, seed, mode):
    import jaxabm_b200 as jx
    import jaxabm_b200.numpy as jnp
    from jaxabm_b200 import random
    from jaxabm_b200.agent import AgentCollection, AgentType
    from jaxabm_b200.model import Model
    Consumer, Producer, ums, cm = user_models.make(jnp, random, AgentType)
    m = Model(params={"price_adjustment_rate": 0.1}, config=jx.ModelConfig(seed=seed, rng_mode=mode),
              update_state_fn=ums, metrics_fn=cm)
    m.add_agent_collection("consumers", AgentCollection(Consumer(), nc))
    m.add_agent_collection("producers", AgentCollection(Producer(), npr))
    for k, v in MARKET_ENV.items():
        m.add_env_state(k, v)
    return m


def oracle_market(nc, npr, seed, mode):
    from oracle import eager, runtime as ort
    Consumer, Producer, ums, cm = user_models.make(eager.jnp, eager.random, eager.AgentTypeBase)
    m = ort.Model(params={"price_adjustment_rate": 0.1}, config=ort.ModelConfig(seed=seed, rng_mode=mode),
                  update_state_fn=eager.wrap_model_fn(ums, mode), metrics_fn=eager.wrap_model_fn(cm, mode, has_key=False))
    m.add_agent_collection("consumers", ort.AgentCollection(eager.wrap_agent_type(Consumer()), nc))
    m.add_agent_collection("producers", ort.AgentCollection(eager.wrap_agent_type(Producer()), npr))
    for k, v in MARKET_ENV.items():
        m.add_env_state(k, v)
    return m


def device_noisy(n, seed, mode):
    import jaxabm_b200 as jx
    import jaxabm_b200.numpy as jnp
    from jaxabm_b200 import random
    from jaxabm_b200.agent import AgentCollection, AgentType
    from jaxabm_b200.model import Model
    Trader, ue, mt = user_models.make_noisy(jnp, random, AgentType)
    m = Model(params={}, config=jx.ModelConfig(seed=seed, rng_mode=mode), update_state_fn=ue, metrics_fn=mt)
    m.add_agent_collection("traders", AgentCollection(Trader(), n))
    for k, v in NOISY_ENV.items():
        m.add_env_state(k, v)
    return m


def oracle_noisy(n, seed, mode):
    from oracle import eager, runtime as ort
    Trader, ue, mt = user_models.make_noisy(eager.jnp, eager.random, eager.AgentTypeBase)
    m = ort.Model(params={}, config=ort.ModelConfig(seed=seed, rng_mode=mode),
                  update_state_fn=eager.wrap_model_fn(ue, mode), metrics_fn=eager.wrap_model_fn(mt, mode, has_key=False))
    m.add_agent_collection("traders", ort.AgentCollection(eager.wrap_agent_type(Trader()), n))
    for k, v in NOISY_ENV.items():
        m.add_env_state(k, v)
    return m
