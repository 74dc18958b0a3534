"""C4-B: households + consumer-goods firms economy -- CPU restatement (TEST INFRASTRUCTURE ONLY,
see ``oracle/__init__.py``).

Follows ``examples/models/advanced_economic_model.py`` of the reference with
``num_capital_firms = num_energy_firms = 0`` and the climate / pandemic modules off (their
defaults are off, ``:1914-1915``):

* ``Household.update``            ``:138-296``   (1 uniform of ``split(key, 4)``)
* ``ConsumerGoodsFirm.update``    ``:389-581``   (1 normal + 1 uniform of ``split(key, 4)``)
* ``update_environment``          ``:1461-1738`` (2 normals of ``split(key, 5)``; INCLUDING the final
  dict comprehension ``:1731-1735`` that writes every pre-existing env entry that is not in its
  exclusion list back over the freshly computed value -- so e.g. ``consumer_goods_price``,
  ``goods_availability``, ``inflation_rate``, ``avg_utility`` keep their initial values forever and
  ``total_income`` freezes at its first-step value.  Observable, therefore restated.)
* ``compute_metrics``             ``:1741-1905`` (31 metrics, Gini from sorted incomes)
* ``create_economy_model``        ``:1909-2232`` (initial env)

dtype rules are JAX's with x64 disabled: columns are float32, Python scalars are weak (scalar
(op) scalar is evaluated in double, scalar (op) array in float32).  Reductions accumulate in
float64 and round once (XLA's float32 reduction order is not restated; tolerance in the tests).

``init_state`` (``:84-136``, ``:329-387``): lognormal / normal / uniform draws follow jax.random;
``random.beta(a=5, b=2)`` is NOT restated (JAX's gamma rejection sampler re-splits keys in a
while loop) -- it is replaced by the distribution-exact construction
``G5 / (G5 + G2)``, ``G_a = -sum_{i<a} log1p(-u_i)``, ``u = uniform(subkey3, (7,))``.
Builder-authored deviation, identical in oracle and engine; parity of initial states vs real JAX
is therefore "unpinned" like every trajectory here.
"""
from __future__ import annotations

import numpy as np

from . import jaxlike as jl
from .runtime import AgentCollection, Model, ModelConfig

f32 = np.float32
i32 = np.int32


def _beta52(keys, mode):
    """Beta(5,2) from 7 uniforms of each key (see module docstring)."""
    n = keys.shape[0]
    u = np.empty((n, 7), dtype=f32)
    for i in range(n):
        u[i] = jl.uniform(keys[i], (7,), mode=mode)
    e = (-np.log1p(-u)).astype(f32)
    g5 = (((e[:, 0] + e[:, 1]) + e[:, 2]) + e[:, 3]) + e[:, 4]
    g2 = e[:, 5] + e[:, 6]
    return (g5 / (g5 + g2)).astype(f32)


def _beta52_fast(keys, mode):
    """Vectorised form of :func:`_beta52` (element j of bits(key,(7,)) for every key)."""
    n = keys.shape[0]
    u = np.empty((n, 7), dtype=f32)
    if mode == 1:
        for j in range(7):
            y0, y1 = jl.threefry2x32(keys[:, 0], keys[:, 1], np.zeros(n, np.uint32), np.full(n, j, np.uint32))
            u[:, j] = jl.bits_to_uniform(y0 ^ y1)
    else:
        # original layout, m = 7 -> padded to 8: x0 = [0..4), x1 = [4,5,6,0]; flat = y0 ++ y1
        for j in range(4):
            x1 = np.full(n, 0 if j == 3 else 4 + j, np.uint32)
            y0, y1 = jl.threefry2x32(keys[:, 0], keys[:, 1], np.full(n, j, np.uint32), x1)
            u[:, j] = jl.bits_to_uniform(y0)
            if j < 3:
                u[:, 4 + j] = jl.bits_to_uniform(y1)
    e = (-np.log1p(-u)).astype(f32)
    g5 = (((e[:, 0] + e[:, 1]) + e[:, 2]) + e[:, 3]) + e[:, 4]
    g2 = e[:, 5] + e[:, 6]
    return (g5 / (g5 + g2)).astype(f32)


class Household:
    """``advanced_economic_model.py:57-296``; 15 state fields."""

    def __init__(self, initial_savings=1000.0, initial_income=100.0, propensity_to_consume=0.8,
                 propensity_to_save=0.1, labor_productivity=1.0, risk_aversion=0.5):
        self.initial_savings = initial_savings
        self.initial_income = initial_income
        self.propensity_to_consume = propensity_to_consume
        self.propensity_to_save = propensity_to_save
        self.labor_productivity = labor_productivity
        self.risk_aversion = risk_aversion

    def init_batch(self, cfg, keys):                                       # :84-136
        m = cfg.rng_mode
        sk = jl.split_batched(keys, 4, m)                                  # [N,4,2]
        n1 = jl.bits_to_normal(jl.random_bits_scalar_batched(sk[:, 0], m))
        savings_factor = np.exp((n1 * f32(0.5)).astype(f32)).astype(f32)  # lognormal(sigma=0.5)
        initial_savings = (f32(self.initial_savings) * savings_factor).astype(f32)
        n2 = jl.bits_to_normal(jl.random_bits_scalar_batched(sk[:, 1], m))
        income_factor = np.maximum(f32(0.3), (n2 * f32(0.2) + f32(1.0)).astype(f32)).astype(f32)
        initial_income = (f32(self.initial_income) * income_factor).astype(f32)
        consume_adj = (_beta52_fast(sk[:, 2], m) * f32(0.4) + f32(0.6)).astype(f32)
        ptc = (f32(self.propensity_to_consume) * consume_adj).astype(f32)
        employed = jl.uniform_scalar_batched(sk[:, 3], mode=m) < f32(0.95)
        n = keys.shape[0]
        return {
            "savings": initial_savings, "income": initial_income,
            "bank_deposits": (initial_savings * f32(0.7)).astype(f32), "cash": (initial_savings * f32(0.3)).astype(f32),
            "debt": 0.0,
            "propensity_to_consume": ptc, "propensity_to_save": self.propensity_to_save,
            "risk_aversion": self.risk_aversion,
            "employed": employed, "productivity": (f32(self.labor_productivity) * income_factor).astype(f32),
            "labor_supply": (employed * f32(1.0)).astype(f32),
            "consumption": 0.0, "utility": 0.0, "taxes_paid": 0.0, "transfers_received": 0.0,
        }

    def update_batch(self, s, model_state, cfg, keys):                     # :138-296
        m = cfg.rng_mode
        env = model_state["env"]
        sk = jl.split_batched(keys, 4, m)
        shock = float(env.get("job_market_condition", 1.0)) * float(env.get("pandemic_impact", 1.0))
        job_loss_prob = f32(0.02 / shock)                                  # scalar / scalar: double, then weak
        job_find_prob = f32(0.1 * shock)
        rv = jl.uniform_scalar_batched(sk[:, 0], mode=m)
        new_employed = np.where(s["employed"], rv > job_loss_prob, rv < job_find_prob)
        labor_supply = (new_employed * f32(1.0)).astype(f32)
        wage_rate = f32(env.get("wage_rate", 1.0))
        tax_rate = f32(env.get("tax_rate", 0.2))
        labor_income = ((labor_supply * s["productivity"]).astype(f32) * wage_rate).astype(f32)
        transfer_rate = f32(env.get("transfer_rate", 0.5))
        unemp = (f32(1.0) - new_employed.astype(f32)).astype(f32)
        transfers = ((unemp * s["income"]).astype(f32) * transfer_rate).astype(f32)
        gross = (labor_income + transfers).astype(f32)
        base_tax = (gross * tax_rate).astype(f32)
        init_inc = f32(self.initial_income)
        prog = np.where(gross > f32(2 * self.initial_income),
                        (f32(0.05) * ((gross / init_inc).astype(f32) - f32(2.0)).astype(f32)).astype(f32), f32(0.0))
        taxes = (base_tax * (f32(1.0) + prog).astype(f32)).astype(f32)
        net = (gross - taxes).astype(f32)
        price_level = f32(env.get("price_level", 1.0))
        goods_av = f32(env.get("goods_availability", 1.0))
        desired = (net * s["propensity_to_consume"]).astype(f32)
        avail = (s["cash"] + (s["bank_deposits"] * f32(0.3)).astype(f32)).astype(f32)
        actual = (np.minimum(desired, avail) * goods_av).astype(f32)
        real_c = (actual / price_level).astype(f32)
        interest_rate = f32(env.get("interest_rate", 0.01))
        interest_income = (s["bank_deposits"] * interest_rate).astype(f32)
        target = (net * s["propensity_to_save"]).astype(f32)
        adj = (target - (s["bank_deposits"] * f32(0.1)).astype(f32)).astype(f32)
        dep = np.minimum(adj, (s["cash"] - actual).astype(f32)).astype(f32)
        pos, neg = np.maximum(f32(0), dep).astype(f32), np.maximum(f32(0), -dep).astype(f32)
        new_cash = (((s["cash"] - actual).astype(f32) - pos).astype(f32) + neg).astype(f32)
        new_dep = ((s["bank_deposits"] + pos).astype(f32) + interest_income).astype(f32)
        cu = np.log1p(real_c).astype(f32)
        su = (s["risk_aversion"] * np.log1p((new_dep / f32(100)).astype(f32)).astype(f32)).astype(f32)
        return {
            "savings": (new_cash + new_dep).astype(f32), "income": gross, "bank_deposits": new_dep,
            "cash": new_cash, "debt": s["debt"],
            "propensity_to_consume": s["propensity_to_consume"], "propensity_to_save": s["propensity_to_save"],
            "risk_aversion": s["risk_aversion"],
            "employed": new_employed, "productivity": s["productivity"], "labor_supply": labor_supply,
            "consumption": real_c, "utility": (cu + su).astype(f32), "taxes_paid": taxes,
            "transfers_received": transfers,
        }


class ConsumerGoodsFirm:
    """``advanced_economic_model.py:299-581``; 19 state fields."""

    def __init__(self, initial_capital=1000.0, initial_cash=500.0, production_efficiency=1.0,
                 labor_elasticity=0.6, capital_elasticity=0.3, energy_elasticity=0.1, markup_rate=0.2):
        self.initial_capital = initial_capital
        self.initial_cash = initial_cash
        self.production_efficiency = production_efficiency
        self.labor_elasticity = labor_elasticity
        self.capital_elasticity = capital_elasticity
        self.energy_elasticity = energy_elasticity
        self.markup_rate = markup_rate

    def init_batch(self, cfg, keys):                                       # :329-387
        m = cfg.rng_mode
        sk = jl.split_batched(keys, 3, m)
        n1 = jl.bits_to_normal(jl.random_bits_scalar_batched(sk[:, 0], m))
        capital = (f32(self.initial_capital) * np.exp((n1 * f32(0.5)).astype(f32)).astype(f32)).astype(f32)
        n2 = jl.bits_to_normal(jl.random_bits_scalar_batched(sk[:, 1], m))
        eff = (f32(self.production_efficiency) *
               np.maximum(f32(0.5), (n2 * f32(0.2) + f32(1.0)).astype(f32)).astype(f32)).astype(f32)
        markup = (f32(self.markup_rate) * (_beta52_fast(sk[:, 2], m) * f32(0.3) + f32(0.1)).astype(f32)).astype(f32)
        cap_pow = np.power(capital, f32(self.capital_elasticity)).astype(f32)
        return {
            "capital_stock": capital, "production_capacity": (eff * cap_pow).astype(f32), "inventory": 0.0,
            "cash": self.initial_cash, "revenue": 0.0, "profit": 0.0, "debt": 0.0,
            "production_efficiency": eff, "labor_demand": 0.0, "energy_usage": 0.0, "goods_produced": 0.0,
            "goods_sold": 0.0, "price": 1.0,
            "markup_rate": markup, "labor_elasticity": self.labor_elasticity,
            "capital_elasticity": self.capital_elasticity, "energy_elasticity": self.energy_elasticity,
            "age": 0, "is_active": True,
        }

    def update_batch(self, s, model_state, cfg, keys):                     # :389-581
        m = cfg.rng_mode
        env = model_state["env"]
        sk = jl.split_batched(keys, 4, m)
        market_demand = env.get("consumer_goods_demand", 100.0)
        market_price = f32(env.get("consumer_goods_price", 1.0))
        wage = f32(env.get("wage_rate", 1.0))
        e_price = f32(env.get("energy_price", 1.0))
        c_price = f32(env.get("capital_price", 1.0))
        rate = f32(env.get("interest_rate", 0.05))
        climate = f32(env.get("climate_impact", 1.0))
        pandemic = f32(env.get("pandemic_impact", 1.0))
        share = 0.01
        md_share = f32(float(market_demand) * share) if not isinstance(market_demand, np.ndarray) \
            else (f32(market_demand) * f32(share)).astype(f32)
        tp = np.maximum(f32(0), (md_share - s["inventory"]).astype(f32)).astype(f32)
        eff, K = s["production_efficiency"], s["capital_stock"]
        le, ce, ee = s["labor_elasticity"], s["capital_elasticity"], s["energy_elasticity"]
        Kc = np.power(K, ce).astype(f32)
        one_e = np.power(f32(1.0), ee).astype(f32)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            denom1 = ((eff * Kc).astype(f32) * one_e).astype(f32)
            base_labor = np.power((tp / denom1).astype(f32), (f32(1) / le).astype(f32)).astype(f32)
            max_labor = (s["cash"] / wage).astype(f32)
            labor = np.minimum(base_labor, max_labor).astype(f32)
            Ll = np.power(labor, le).astype(f32)
            denom2 = ((eff * Kc).astype(f32) * Ll).astype(f32)
            base_energy = np.power((tp / denom2).astype(f32), (f32(1) / ee).astype(f32)).astype(f32)
            remaining = (s["cash"] - (labor * wage).astype(f32)).astype(f32)
            max_energy = (remaining / e_price).astype(f32)
            energy = np.minimum(base_energy, max_energy).astype(f32)
            production = (((eff * Ll).astype(f32) * Kc).astype(f32) * np.power(energy, ee).astype(f32)).astype(f32)
            affected = ((production * climate).astype(f32) * pandemic).astype(f32)
            noise = (jl.bits_to_normal(jl.random_bits_scalar_batched(sk[:, 0], m)) * f32(0.05) + f32(1.0)).astype(f32)
            actual_prod = (affected * noise).astype(f32)
            new_inv = (s["inventory"] + actual_prod).astype(f32)
            cost = (((labor * wage).astype(f32) + (energy * e_price).astype(f32)).astype(f32) +
                    ((K * c_price).astype(f32) * f32(0.05)).astype(f32)).astype(f32)
            one_mk = (f32(1) + s["markup_rate"]).astype(f32)
            unit_cost = np.where(actual_prod > f32(0), (cost / actual_prod).astype(f32),
                                 (s["price"] / one_mk).astype(f32)).astype(f32)
            target_price = (unit_cost * one_mk).astype(f32)
            speed = 0.3
            new_price = ((s["price"] * f32(1 - speed)).astype(f32) + (target_price * f32(speed)).astype(f32)).astype(f32)
            pressure = (s["inventory"] / (actual_prod + f32(0.1)).astype(f32)).astype(f32)
            discount = np.maximum(f32(0), np.minimum(f32(0.2), (f32(0.05) * pressure).astype(f32))).astype(f32)
            final_price = (new_price * (f32(1) - discount).astype(f32)).astype(f32)
            competitiveness = np.power((market_price / final_price).astype(f32), f32(1.2)).astype(f32)
            sales_rand = (jl.uniform_scalar_batched(sk[:, 1], mode=m) * f32(0.4) + f32(0.8)).astype(f32)
            potential = ((md_share * competitiveness).astype(f32) * sales_rand).astype(f32)
            sales = np.minimum(potential, new_inv).astype(f32)
            final_inv = (new_inv - sales).astype(f32)
            revenue = (sales * final_price).astype(f32)
            total_costs = (cost + (s["debt"] * rate).astype(f32)).astype(f32)
            profit = (revenue - total_costs).astype(f32)
            new_cash = (((s["cash"] + revenue).astype(f32) - (labor * wage).astype(f32)).astype(f32) -
                        (energy * e_price).astype(f32)).astype(f32)
            inv_ratio = np.where(profit > f32(0), f32(0.3), f32(0.0)).astype(f32)
            investment = (profit * inv_ratio).astype(f32)
            actual_inv = np.minimum(investment, (new_cash * f32(0.5)).astype(f32)).astype(f32)
            purchases = (actual_inv / c_price).astype(f32)
            new_K = ((K * f32(1 - 0.05)).astype(f32) + purchases).astype(f32)
            final_cash = (new_cash - actual_inv).astype(f32)
            viable = (final_cash > f32(0)) & (new_K > f32(0))
        return {
            "capital_stock": new_K, "production_capacity": (new_K * eff).astype(f32), "inventory": final_inv,
            "cash": final_cash, "revenue": revenue, "profit": profit, "debt": s["debt"],
            "production_efficiency": eff, "labor_demand": labor, "energy_usage": energy,
            "goods_produced": actual_prod, "goods_sold": sales, "price": final_price,
            "markup_rate": s["markup_rate"], "labor_elasticity": le, "capital_elasticity": ce,
            "energy_elasticity": ee, "age": (s["age"] + i32(1)).astype(i32), "is_active": viable,
        }


def _fsum(a):
    return f32(np.sum(np.asarray(a, dtype=np.float64)))


def _fmean(a):
    a = np.asarray(a)
    return f32(f32(np.sum(a.astype(np.float64))) / f32(a.shape[0]))


_EXCLUDED = ("time_step", "wage_rate", "price_level", "interest_rate", "gdp", "gdp_growth", "total_labor_supply",
             "total_labor_demand", "employment_rate", "unemployment_rate", "climate_impact", "pandemic_impact")


def update_environment(env_state, agent_states, params, key, mode=None):   # :1461-1738
    sk = jl.split(key, 5, mode)
    hh = agent_states.get("households", {})
    if not hh:
        return env_state
    total_labor_supply = _fsum(hh["labor_supply"])
    total_consumption = _fsum(hh["consumption"])
    total_savings = _fsum(hh["savings"])
    total_bank_deposits = _fsum(hh["bank_deposits"])
    total_income = _fsum(hh["income"])
    employment_rate = _fmean(hh["employed"].astype(f32))
    avg_utility = _fmean(hh["utility"])
    cf = agent_states.get("consumer_firms", {})
    if cf:
        produced = _fsum(cf["goods_produced"]); sold = _fsum(cf["goods_sold"]); inventory = _fsum(cf["inventory"])
        cg_price = _fmean(cf["price"]); cf_labor = _fsum(cf["labor_demand"]); cf_energy = _fsum(cf["energy_usage"])
        cf_profit = _fsum(cf["profit"])
        replacement = f32(_fsum(cf["capital_stock"]) * f32(0.05))
        expansion = f32(np.maximum(f32(0.0), cf_profit) * f32(0.3))
        cf_investment = f32(replacement + expansion)
    else:
        produced = sold = inventory = 0.0
        cg_price = 1.0
        cf_labor = cf_energy = cf_investment = 0.0
    capital_goods_produced, capital_goods_price, capital_labor = 0.0, 2.0, 0.0
    energy_price, energy_labor = 1.0, 0.0
    tax_revenue = govt_spending = public_debt = 0.0
    total_deposits, total_loans = total_bank_deposits, 0.0
    old_wage = f32(env_state.get("wage_rate", 1.0))
    old_price = f32(env_state.get("price_level", 1.0))
    old_rate = f32(env_state.get("interest_rate", 0.05))
    total_labor_demand = f32(f32(f32(cf_labor) + f32(capital_labor)) + f32(energy_labor))
    tightness = f32(total_labor_demand / f32(total_labor_supply + f32(1e-6))) if total_labor_supply > f32(1e-6) else f32(1.0)
    wage_pressure = f32(f32(tightness - f32(1.0)) * f32(0.2))
    wage_noise = f32(jl.normal(sk[0], (), mode) * f32(0.01))
    wage_change = f32(np.clip(f32(wage_pressure + wage_noise), f32(-0.05), f32(0.05)))
    new_wage = f32(np.maximum(f32(0.1), f32(old_wage * f32(f32(1.0) + wage_change))))
    w = np.array([0.6, 0.3, 0.1], dtype=f32)
    p = np.array([cg_price, capital_goods_price, energy_price], dtype=f32)
    prod = (w * p).astype(f32)
    new_price = f32(np.maximum(f32(0.1), f32(f32(prod[0] + prod[1]) + prod[2])))
    inflation = f32(f32(new_price / np.maximum(f32(0.1), old_price)) - f32(1.0))
    unemployment = f32(f32(1.0) - employment_rate)
    inflation_gap = f32(inflation - f32(0.02))
    output_gap = f32(f32(-0.5) * f32(unemployment - f32(0.05)))
    taylor = f32(f32(f32(0.02) + f32(f32(1.5) * inflation_gap)) + f32(f32(0.5) * output_gap))
    change = f32(f32(taylor - old_rate) * f32(0.3))
    rate_noise = f32(jl.normal(sk[1], (), mode) * f32(0.005))
    change = f32(np.clip(f32(change + rate_noise), f32(-0.02), f32(0.02)))
    new_rate = f32(np.maximum(f32(0.01), f32(old_rate + change)))
    gdp = f32(f32(f32(produced) * f32(cg_price)) + f32(capital_goods_produced * capital_goods_price))
    gdp = f32(gdp + f32(govt_spending))
    gdp = f32(np.maximum(f32(0.1), gdp))
    prev_gdp = f32(np.maximum(f32(0.1), f32(env_state.get("gdp", gdp))))
    gdp_growth = f32(f32(gdp / prev_gdp) - f32(1.0))
    household_count = 1                                    # households have no 'position' -> jnp.array([0]).shape[0]
    income_per_capita = f32(total_income / f32(max(1, household_count)))
    new = {
        "time_step": env_state.get("time_step", 0) + 1,
        "wage_rate": new_wage, "total_labor_supply": total_labor_supply, "total_labor_demand": total_labor_demand,
        "employment_rate": employment_rate, "unemployment_rate": f32(f32(1.0) - employment_rate),
        "job_market_condition": tightness,
        "price_level": new_price, "inflation_rate": inflation, "consumer_goods_price": cg_price,
        "capital_goods_price": capital_goods_price, "energy_price": energy_price,
        "consumer_goods_supply": produced, "consumer_goods_demand": total_consumption,
        "consumer_goods_inventory": inventory,
        "goods_availability": f32(np.minimum(f32(1.0), f32(f32(produced) / np.maximum(f32(1e-5), total_consumption)))),
        "interest_rate": new_rate, "total_savings": total_savings, "total_deposits": total_deposits,
        "total_loans": total_loans,
        "tax_revenue": tax_revenue, "govt_spending": govt_spending, "public_debt": public_debt,
        "debt_to_gdp": f32(f32(public_debt) / np.maximum(f32(0.1), gdp)),
        "gdp": gdp, "gdp_growth": gdp_growth, "avg_utility": avg_utility, "total_income": total_income,
        "income_per_capita": income_per_capita,
        "climate_impact": 1.0, "pandemic_impact": 1.0,
        "overall_climate_impact": 1.0, "overall_pandemic_impact": 1.0,
    }
    for k, v in env_state.items():                         # :1731-1735 -- old values win outside the list
        if k not in _EXCLUDED:
            new[k] = v
    return new


def gini_sorted(incomes):
    """``:1792-1809``: index * sorted incomes in float32, sums rounded once."""
    x = np.sort(np.asarray(incomes, dtype=f32))
    n = x.shape[0]
    idx = np.arange(1, n + 1, dtype=i32)
    income_sum = _fsum(x)
    if not income_sum > f32(1e-6):
        return f32(0.0)
    weighted = _fsum((idx.astype(f32) * x).astype(f32))
    a = f32(f32(f32(2) * weighted) / f32(f32(n) * income_sum))
    b = f32((n + 1) / n)
    return f32(a - b)


METRIC_NAMES = ["gdp", "gdp_growth", "inflation", "unemployment", "wage_rate", "interest_rate",
                "goods_availability", "labor_market_tightness", "consumer_price", "capital_price", "energy_price",
                "utility", "income_per_capita", "inequality", "consumer_sector_share", "capital_sector_share",
                "energy_sector_share", "govt_sector_share", "technology_level", "capital_investment", "energy_demand",
                "energy_supply", "renewable_share", "carbon_emissions", "sustainability_index", "debt_to_gdp",
                "climate_impact", "pandemic_impact", "economic_health"]


def compute_metrics(env, agent_states, params):            # :1741-1905
    g = lambda k, d: env.get(k, d)
    gdp, gdp_growth = g("gdp", 0.1), g("gdp_growth", 0.0)
    inflation = g("inflation_rate", 0.0)
    unemployment, wage = g("unemployment_rate", 0.0), g("wage_rate", 1.0)
    tightness = g("job_market_condition", 1.0)
    goods_av, c_price = g("goods_availability", 1.0), g("consumer_goods_price", 1.0)
    k_price, e_price = g("capital_goods_price", 2.0), g("energy_price", 1.0)
    tech = g("avg_technology_level", 1.0)
    k_demand = g("capital_goods_demand", 0.0)
    e_demand, e_supply = g("energy_demand", 0.0), g("energy_supply", 0.0)
    carbon, renew = g("carbon_emissions", 0.0), g("avg_renewable_fraction", 0.2)
    rate, debt_gdp = g("interest_rate", 0.05), g("debt_to_gdp", 0.0)
    avg_utility, ipc = g("avg_utility", 0.0), g("income_per_capita", 0.0)
    hh = agent_states.get("households", {})
    gini = f32(0.0)
    if hh and "income" in hh:
        gini = gini_sorted(hh["income"])
    climate, pandemic = g("climate_impact", 1.0), g("pandemic_impact", 1.0)
    c_gdp = g("consumer_goods_sold", 0.0) * c_price        # never set by update_environment -> 0.0
    k_gdp = g("capital_goods_sold", 0.0) * k_price
    e_gdp = g("energy_sold", 0.0) * e_price
    g_gdp = g("govt_spending", 0.0)
    total = c_gdp + k_gdp + e_gdp + g_gdp
    share = lambda v: f32((v / total) * 100) if total > 0.1 else f32(0.0)
    health = f32(f32(0.25) * f32(f32(1.0) - f32(unemployment)))
    health = f32(health + f32(f32(0.15) * f32(np.clip(f32(f32(gdp_growth) * f32(10)), f32(-1.0), f32(1.0)))))
    health = f32(health + f32(f32(0.15) * f32(f32(1.0) - f32(abs(f32(f32(inflation) - f32(0.02))) * f32(10)))))
    health = f32(health + f32(f32(0.10) * f32(f32(1.0) - f32(min(f32(debt_gdp), f32(1.0))))))
    health = f32(health + f32(f32(0.10) * f32(goods_av)))
    health = f32(health + f32(f32(f32(0.10) * f32(avg_utility)) / f32(max(0.1, 2.0))))
    health = f32(health + f32(f32(0.05) * f32(f32(1.0) - f32(min(gini, f32(1.0))))))
    health = f32(health + f32(f32(f32(0.05) * f32(tech)) / f32(2.0)))
    health = f32(health + f32(f32(0.05) * f32(renew)))
    health_index = f32(np.clip(f32(health * f32(100)), f32(0), f32(100)))
    sustain = (0.4 * renew + 0.3 * (1.0 - min(carbon / 100.0, 1.0)) + 0.2 * float(np.clip((tech - 1.0) * 0.5, 0.0, 1.0)) +
               0.1 * (1.0 - max(0.0, (climate - 1.0)))) * 100
    nn = lambda v, nan: f32(nan) if np.isnan(f32(v)) else f32(v)
    return {
        "gdp": nn(gdp, 0.1), "gdp_growth": nn(f32(f32(gdp_growth) * f32(100)), 0.0),
        "inflation": nn(f32(f32(inflation) * f32(100)), 0.0), "unemployment": nn(f32(f32(unemployment) * f32(100)), 0.0),
        "wage_rate": nn(wage, 1.0), "interest_rate": nn(f32(f32(rate) * f32(100)), 0.0),
        "goods_availability": nn(f32(f32(goods_av) * f32(100)), 100.0), "labor_market_tightness": nn(tightness, 1.0),
        "consumer_price": nn(c_price, 1.0), "capital_price": nn(k_price, 2.0), "energy_price": nn(e_price, 1.0),
        "utility": nn(avg_utility, 0.0), "income_per_capita": nn(ipc, 0.0), "inequality": nn(gini, 0.0),
        "consumer_sector_share": nn(share(c_gdp), 0.0), "capital_sector_share": nn(share(k_gdp), 0.0),
        "energy_sector_share": nn(share(e_gdp), 0.0), "govt_sector_share": nn(share(g_gdp), 0.0),
        "technology_level": nn(tech, 1.0), "capital_investment": nn(k_demand, 0.0),
        "energy_demand": nn(e_demand, 0.0), "energy_supply": nn(e_supply, 0.0),
        "renewable_share": nn(renew * 100, 20.0), "carbon_emissions": nn(carbon, 0.0),
        "sustainability_index": nn(sustain, 50.0), "debt_to_gdp": nn(f32(f32(debt_gdp) * f32(100)), 0.0),
        "climate_impact": nn((1.0 - (climate - 1.0)) * 100, 0.0), "pandemic_impact": nn((1.0 - pandemic) * 100, 0.0),
        "economic_health": nn(health_index, 50.0),
    }


def initial_env(num_households, num_consumer_firms, tax_rate=0.2, interest_rate=0.05, energy_price=1.0,
                wage_rate=1.0, initial_income=100.0, initial_savings=1000.0):
    """``create_economy_model`` env block (``:2129-2230``) with no capital / energy firms."""
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


def create_economy_model(num_households=1000, num_consumer_firms=50, tax_rate=0.2, interest_rate=0.05,
                         energy_price=1.0, wage_rate=1.0, household_params=None, consumer_firm_params=None,
                         seed=42, params=None, config=None):
    """``:1909-2232`` with ``num_capital_firms = num_energy_firms = 0``."""
    if params is not None:
        tax_rate = params.get("tax_rate", tax_rate)
        interest_rate = params.get("interest_rate", interest_rate)
        energy_price = params.get("energy_price", energy_price)
        num_households = params.get("num_households", num_households)
        num_consumer_firms = params.get("num_consumer_firms", num_consumer_firms)
    if config is None:
        config = ModelConfig(seed=seed, steps=100)
    mode = config.rng_mode
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
    model = Model(params=dict(params or {}), config=config,
                  update_state_fn=lambda e, a, p, k: update_environment(e, a, p, k, mode),
                  metrics_fn=compute_metrics)
    model.add_agent_collection("households", AgentCollection(Household(**hd), num_households))
    if num_consumer_firms > 0:
        model.add_agent_collection("consumer_firms", AgentCollection(ConsumerGoodsFirm(**fd), num_consumer_firms))
    for k, v in initial_env(num_households, num_consumer_firms, tax_rate, interest_rate, energy_price, wage_rate,
                            hd["initial_income"], hd["initial_savings"]).items():
        model.add_env_state(k, v)
    return model
