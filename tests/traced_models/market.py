"""The consumer / producer economy of the reference's integration test
(``tests/integration/test_integration.py:20-183``) written as PLAIN user code against the
``jnp`` / ``random`` names a backend provides -- no registered kernels.  ``make(jnp, random, ...)``
builds the classes for a backend: ``jaxabm_b200.numpy`` / ``jaxabm_b200.random`` (traced into a
generated CUDA kernel) or the NumPy batch backend of the oracle (``oracle/eager.py``)."""


def make(jnp, random, AgentType):
    class Consumer(AgentType):
        def __init__(self, base_income=1.0, propensity_to_consume=0.8):
            self.base_income = base_income
            self.propensity_to_consume = propensity_to_consume

        def init_state(self, model_config, key):
            income = self.base_income * (0.8 + 0.4 * random.uniform(key))
            return {"savings": 0.0, "consumption": 0.0, "utility": 0.0, "income": income}

        def update(self, state, model_state, model_config, key):
            price_level = model_state["env"].get("price_level", 1.0)
            consumption = self.propensity_to_consume * state["income"] / price_level
            savings = state["savings"] + (state["income"] - consumption * price_level)
            utility = jnp.log(consumption + 1.0)
            return {"savings": savings, "consumption": consumption, "utility": utility, "income": state["income"]}

    class Producer(AgentType):
        def __init__(self, initial_capital=10.0, productivity=1.0, reinvestment_rate=0.3):
            self.initial_capital = initial_capital
            self.productivity = productivity
            self.reinvestment_rate = reinvestment_rate

        def init_state(self, model_config, key):
            capital = self.initial_capital * (0.8 + 0.4 * random.uniform(key))
            return {"capital": capital, "production": 0.0, "profit": 0.0}

        def update(self, state, model_state, model_config, key):
            price_level = model_state["env"].get("price_level", 1.0)
            production = self.productivity * state["capital"] ** 0.7
            revenue = production * price_level
            costs = 0.1 * state["capital"] + 0.05 * production
            profit = revenue - costs
            capital = state["capital"] + self.reinvestment_rate * profit
            return {"capital": capital, "production": production, "profit": profit}

    def update_model_state(env_state, agent_states, model_params, key):
        total_consumption = jnp.sum(agent_states["consumers"]["consumption"])
        total_production = jnp.sum(agent_states["producers"]["production"])
        price_level = env_state.get("price_level", 1.0)
        price_adjustment_rate = model_params.get("price_adjustment_rate", 0.1)
        production_consumption_ratio = (total_production + 1e-8) / (total_consumption + 1e-8)
        price_change = price_adjustment_rate * (1.0 - production_consumption_ratio)
        new_price_level = price_level * (1.0 + price_change)
        new_price_level = jnp.maximum(0.5, jnp.minimum(2.0, new_price_level))
        gdp = total_production * new_price_level
        unemployment = jnp.maximum(0.0, jnp.minimum(0.5, 1.0 - production_consumption_ratio))
        new_env_state = dict(env_state)
        new_env_state["price_level"] = new_price_level
        new_env_state["gdp"] = gdp
        new_env_state["unemployment"] = unemployment
        new_env_state["total_consumption"] = total_consumption
        new_env_state["total_production"] = total_production
        return new_env_state

    def compute_metrics(env_state, agent_states, model_params):
        return {"gdp": env_state.get("gdp", 0.0), "price_level": env_state.get("price_level", 1.0),
                "unemployment": env_state.get("unemployment", 0.0),
                "avg_utility": jnp.mean(agent_states["consumers"]["utility"]),
                "avg_profit": jnp.mean(agent_states["producers"]["profit"])}

    return Consumer, Producer, update_model_state, compute_metrics


def make_noisy(jnp, random, AgentType):
    """A model no registered kernel covers: per-agent draws in update, int and bool state, env-level noise."""
    class Trader(AgentType):
        def init_state(self, model_config, key):
            k1, k2 = random.split(key)
            return {"wealth": 10.0 + 5.0 * random.normal(k1), "active": random.uniform(k2) < 0.9, "trades": 0}

        def update(self, state, model_state, model_config, key):
            k1, k2, k3 = random.split(key, 3)
            shock = model_state["env"]["volatility"] * random.normal(k1)
            gain = jnp.where(state["active"], state["wealth"] * (0.01 + shock), 0.0)
            wealth = jnp.maximum(0.0, state["wealth"] + gain)
            quits = random.uniform(k2) < 0.02
            joins = random.uniform(k3) < 0.10
            active = jnp.where(state["active"], ~quits, joins)
            trades = state["trades"] + jnp.where(state["active"], 1, 0)
            return {"wealth": wealth, "active": active, "trades": trades}

    def update_env(env_state, agent_states, params, key):
        a = agent_states["traders"]
        participation = jnp.mean(a["active"].astype(float))
        new = dict(env_state)
        new["volatility"] = jnp.clip(env_state["volatility"] * (1.0 + 0.1 * random.normal(key)), 0.001, 0.2)
        new["participation"] = participation
        new["steps_done"] = env_state["steps_done"] + 1
        return new

    def metrics(env_state, agent_states, params):
        a = agent_states["traders"]
        return {"mean_wealth": jnp.mean(a["wealth"]), "max_wealth": jnp.max(a["wealth"]), "min_wealth": jnp.min(a["wealth"]),
                "n_active": jnp.sum(a["active"]), "total_trades": jnp.sum(a["trades"]),
                "volatility": env_state["volatility"], "participation": env_state["participation"],
                "steps_done": env_state["steps_done"],
                "rich_share": jnp.sum(jnp.where(a["wealth"] > 12.0, a["wealth"], 0.0)) / jnp.sum(a["wealth"])}

    return Trader, update_env, metrics
