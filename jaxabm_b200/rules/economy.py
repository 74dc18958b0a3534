"""Households + consumer-goods firms economy (C4-B): host handles for ``rule_household`` /
``rule_firm`` and the ``JXB_PROGRAM_ECONOMY`` tail (``csrc/economy.cuh``).

Mirrors ``examples/models/advanced_economic_model.py`` of the reference: ``Household`` (:57-296),
``ConsumerGoodsFirm`` (:299-581), ``update_environment`` (:1461-1738), ``compute_metrics``
(:1741-1905), ``create_economy_model`` (:1909-2232) -- with ``num_capital_firms =
num_energy_firms = 0`` (the factory skips a type whose count is 0, :2060, :2090) and the climate /
pandemic modules off (their defaults, :1914-1915).  Capital-goods / energy firms and the two
modules are not registered device rules: asking for them raises ``UnregisteredRuleError``.

``init_state`` runs on the device from the per-agent keys; ``random.beta`` is replaced by a
distribution-exact construction from 7 uniforms (see ``csrc/economy.cuh::beta52``), every other
draw follows ``jax.random``.
"""
from __future__ import annotations

from ..agent import AgentCollection, AgentType, UnregisteredRuleError
from ..core import ModelConfig
from ..model import Model
from . import program


class Household(AgentType):
    """15 state fields (``:113-136``)."""
    jxb_rule = "household"

    def __init__(self, initial_savings=1000.0, initial_income=100.0, propensity_to_consume=0.8,
                 propensity_to_save=0.1, labor_productivity=1.0, risk_aversion=0.5):
        self.initial_savings = initial_savings
        self.initial_income = initial_income
        self.propensity_to_consume = propensity_to_consume
        self.propensity_to_save = propensity_to_save
        self.labor_productivity = labor_productivity
        self.risk_aversion = risk_aversion

    def jxb_params(self):
        return [self.initial_savings, self.initial_income, self.propensity_to_consume, self.propensity_to_save,
                self.labor_productivity, self.risk_aversion]


class ConsumerGoodsFirm(AgentType):
    """19 state fields (``:358-387``)."""
    jxb_rule = "consumer_firm"

    def __init__(self, initial_capital=1000.0, initial_cash=500.0, production_efficiency=1.0, labor_elasticity=0.6,
                 capital_elasticity=0.3, energy_elasticity=0.1, markup_rate=0.2):
        self.initial_capital = initial_capital
        self.initial_cash = initial_cash
        self.production_efficiency = production_efficiency
        self.labor_elasticity = labor_elasticity
        self.capital_elasticity = capital_elasticity
        self.energy_elasticity = energy_elasticity
        self.markup_rate = markup_rate

    def jxb_params(self):
        return [self.initial_capital, self.initial_cash, self.production_efficiency, self.labor_elasticity,
                self.capital_elasticity, self.energy_elasticity, self.markup_rate]


@program("economy")
def update_environment(env_state, agent_states, params, key):
    """Runs on the device as the tail of ``economy_step_kernel``."""
    raise RuntimeError("device-resident model function; it is not called on the host")


@program("economy")
def compute_metrics(env_state, agent_states, params):
    """Runs on the device as the tail of ``gini_accumulate_kernel``."""
    raise RuntimeError("device-resident model function; it is not called on the host")


def initial_env(num_households, num_consumer_firms, tax_rate=0.2, interest_rate=0.05, energy_price=1.0,
                wage_rate=1.0, initial_income=100.0, initial_savings=1000.0):
    """The ``add_env_state`` block of ``create_economy_model`` (``:2129-2230``), no capital / energy firms."""
    initial_gdp = num_households * initial_income * 0.8
    cons_prod = num_consumer_firms * 20.0
    cap_prod = 0 * 10.0
    en_prod = 0 * 50.0
    total_savings = num_households * initial_savings
    return {
        "time_step": 0, "wage_rate": wage_rate, "price_level": 1.0, "interest_rate": interest_rate,
        "tax_rate": tax_rate, "energy_price": energy_price, "fossil_fuel_price": 0.8,
        "climate_policy_strength": 0.2, "carbon_price": 0.1, "renewable_subsidy": 0.05,
        "gdp": initial_gdp, "inflation_rate": 0.02,
        "job_market_condition": 1.0, "employment_rate": 0.95, "unemployment_rate": 0.05,
        "total_labor_supply": num_households * 0.95, "total_labor_demand": num_households * 0.95,
        "consumer_goods_price": 1.0, "consumer_goods_supply": cons_prod, "consumer_goods_demand": cons_prod * 0.9,
        "consumer_goods_inventory": cons_prod * 0.1,
        "capital_goods_price": 2.0, "capital_goods_supply": cap_prod, "capital_goods_demand": cap_prod * 0.8,
        "capital_goods_inventory": cap_prod * 0.2,
        "energy_supply": en_prod, "household_energy_demand": en_prod * 0.3,
        "consumer_firms_energy_usage": en_prod * 0.4, "capital_firms_energy_usage": en_prod * 0.3,
        "goods_availability": 1.0,
        "consumer_firms_investment": cap_prod * 0.5, "energy_firms_investment": cap_prod * 0.3,
        "total_savings": total_savings, "total_deposits": total_savings * 0.7, "total_loans": total_savings * 0.5,
        "tax_revenue": initial_gdp * tax_rate, "govt_spending": initial_gdp * tax_rate * 1.1,
        "public_debt": initial_gdp * 0.6, "debt_to_gdp": 0.6,
        "avg_utility": 1.0, "income_per_capita": initial_income,
        "climate_impact": 1.0, "pandemic_impact": 1.0, "pandemic_infected_rate": 0.0,
        "climate_trend": 0.0, "extreme_event_magnitude": 0.0,
    }


def create_economy_model(num_households=1000, num_consumer_firms=50, num_capital_firms=0, num_energy_firms=0,
                         enable_climate_module=False, enable_pandemic_module=False, tax_rate=0.2,
                         interest_rate=0.05, energy_price=1.0, wage_rate=1.0, household_params=None,
                         consumer_firm_params=None, seed=42, params=None, config=None):
    """``advanced_economic_model.py:1909-2232``.  NOTE the reference defaults ``num_capital_firms=20,
    num_energy_firms=10``; those agent types have no registered kernel, so the defaults here are 0."""
    if params is not None:
        tax_rate = params.get("tax_rate", tax_rate)
        interest_rate = params.get("interest_rate", interest_rate)
        energy_price = params.get("energy_price", energy_price)
        num_households = params.get("num_households", num_households)
        num_consumer_firms = params.get("num_consumer_firms", num_consumer_firms)
        num_capital_firms = params.get("num_capital_firms", num_capital_firms)
        num_energy_firms = params.get("num_energy_firms", num_energy_firms)
    if num_capital_firms or num_energy_firms:
        raise UnregisteredRuleError("CapitalGoodsFirm / EnergyFirm have no registered CUDA rule; "
                                    "use num_capital_firms=0, num_energy_firms=0")
    if enable_climate_module or enable_pandemic_module:
        raise UnregisteredRuleError("the climate / pandemic modules are not registered device programs")
    if config is None:
        config = ModelConfig(seed=seed, steps=100, track_history=True, collect_interval=1)
    hd = {"initial_savings": 1000.0, "initial_income": 100.0, "propensity_to_consume": 0.8, "propensity_to_save": 0.1}
    hd.update(household_params or {})
    fd = {"initial_capital": 1000.0, "initial_cash": 500.0, "production_efficiency": 1.0, "markup_rate": 0.2}
    fd.update(consumer_firm_params or {})
    if params:
        for k in ("propensity_to_consume", "propensity_to_save"):
            if k in params:
                hd[k] = params[k]
        for k in ("production_efficiency", "markup_rate"):
            if k in params:
                fd[k] = params[k]
    model = Model(params={"enable_climate_module": False, "enable_pandemic_module": False, "tax_rate": tax_rate,
                          "interest_rate": interest_rate, "energy_price": energy_price, "wage_rate": wage_rate,
                          "num_households": num_households, "num_consumer_firms": num_consumer_firms,
                          "num_capital_firms": 0, "num_energy_firms": 0, **(params or {})},
                  config=config, update_state_fn=update_environment, metrics_fn=compute_metrics)
    model.add_agent_collection("households", AgentCollection(Household(**hd), num_households))
    if num_consumer_firms > 0:
        model.add_agent_collection("consumer_firms", AgentCollection(ConsumerGoodsFirm(**fd), num_consumer_firms))
    for k, v in initial_env(num_households, num_consumer_firms, tax_rate, interest_rate, energy_price, wage_rate,
                            hd["initial_income"], hd["initial_savings"]).items():
        model.add_env_state(k, v)
    return model
