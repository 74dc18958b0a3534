"""Workload rules, restated for the CPU oracle (NumPy, float32/int32 as JAX x64-off).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference-defined rules (restated, batch form):

* C1 random walk      -- ``examples/basic_example.py:20-182``
* C4-A market         -- ``tests/integration/test_integration.py:20-183``
* C5 sweep model      -- ``tests/unit/test_analysis.py:22-144``
* contract models     -- ``tests/unit/test_agent.py:44-100``, ``tests/unit/test_model.py:20-48``,
                         ``tests/conftest.py:55-110``

Builder-authored rules (the reference defines only their state layout; parity for
these is *unpinned*; the definitions are in DESIGN.md):

* C2 Schelling        -- layout ``examples/models/schelling_model.py:26-31,119-139``
* C3 SIR on a network -- layout ``jaxabm/agentpy.py:557,574-582``
"""
from __future__ import annotations

from typing import Any, Dict

import numpy as np

from . import jaxlike as jl
from . import facade
from .runtime import AgentCollection, Model, ModelConfig, unbatched

f32 = np.float32
i32 = np.int32


# ===========================================================================
# C1  random walk  (examples/basic_example.py)
# ===========================================================================

class RandomWalker(facade.Agent):
    def setup(self):                                             # basic_example.py:23-31
        return {
            "position": np.array([0.5, 0.5], dtype=f32),
            "velocity": np.array([0.01, 0.01], dtype=f32),
            "color": 0,
            "steps_taken": 0,
        }

    def step_batch(self, s, model_state):                        # basic_example.py:33-69
        position, velocity, color = s["position"], s["velocity"], s["color"]
        steps_taken = (s["steps_taken"] + i32(1)).astype(i32)
        env = model_state.get("env", {})
        bounds = np.asarray(env.get("bounds", np.array([0.0, 1.0])), dtype=f32)
        new_position = (position + velocity).astype(f32)
        x_bounce = (new_position[:, 0] <= bounds[0]) | (new_position[:, 0] >= bounds[1])
        y_bounce = (new_position[:, 1] <= bounds[0]) | (new_position[:, 1] >= bounds[1])
        sign = np.stack([1 - 2 * x_bounce.astype(i32), 1 - 2 * y_bounce.astype(i32)], axis=1)
        new_velocity = (velocity * sign.astype(f32)).astype(f32)
        new_position = np.clip(new_position, bounds[0], bounds[1]).astype(f32)
        any_bounce = np.logical_or(x_bounce, y_bounce)
        new_color = np.where(any_bounce, 1 - color, color).astype(i32)
        return {"position": new_position, "velocity": new_velocity,
                "color": new_color, "steps_taken": steps_taken}


class RandomWalkModel(facade.Model):
    """``basic_example.py:72-182``.  ``step()`` touches only ``_jax_model`` env and then
    ``record_data()`` returns early (``state['agents']`` never exists, F11); the facade's
    overlay then restores ``Environment.state`` -- so only the distance metrics move."""

    def setup(self):
        n = self.p.get("n_agents", 50)
        self.walkers = self.add_agents(n, RandomWalker)
        self.env.add_state("bounds", np.array([0.0, 1.0], dtype=f32))
        self.env.add_state("time", 0)
        self.env.add_state("mean_x", 0.5)
        self.env.add_state("mean_y", 0.5)
        self.env.add_state("num_red", n)
        self.env.add_state("num_blue", 0)

    def step(self):                                              # basic_example.py:91-100
        jm = self._jax_model
        if jm is not None and jm.state and "env" in jm.state:
            t = jm.state["env"].get("time", 0)
            jm.add_env_state("time", t + 1)                      # clobbered by the overlay

    def compute_metrics(self, env_state, agent_states, model_params):   # basic_example.py:141-182
        out = {k: env_state.get(k, d) for k, d in
               (("mean_x", 0.5), ("mean_y", 0.5), ("num_red", 0), ("num_blue", 0), ("time", 0))}
        # basic_example.py:151 looks up 'walkers' but the collection is auto-named
        # 'randomwalkers' (agentpy.py:960-961) -> the default branch is what runs.
        if "walkers" in agent_states and "position" in agent_states["walkers"]:
            d = walker_distances(agent_states["walkers"]["position"])
            out["mean_distance"] = f32(np.mean(d))
            out["max_distance"] = f32(np.max(d))
        else:
            out["mean_distance"] = 0.0
            out["max_distance"] = 0.0
        return out


def walker_distances(position):
    c = position - np.array([0.5, 0.5], dtype=f32)
    return np.sqrt((c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1]).astype(f32)).astype(f32)


class RandomWalkModelNamed(RandomWalkModel):
    """Same model with the collection registered under the name the metrics expect
    (``add_agents(..., name='walkers')``): the distance metrics become live."""

    def setup(self):
        n = self.p.get("n_agents", 50)
        self.walkers = self.add_agents(n, RandomWalker, name="walkers")
        self.env.add_state("bounds", np.array([0.0, 1.0], dtype=f32))
        self.env.add_state("time", 0)
        self.env.add_state("mean_x", 0.5)
        self.env.add_state("mean_y", 0.5)
        self.env.add_state("num_red", n)
        self.env.add_state("num_blue", 0)


class ScaledRandomWalker(RandomWalker):
    """Roofline variant (SURVEY 8d, C1 scaled): per-agent start ``uniform(key_i)`` on both
    axes and velocity ``0.01*(2u-1)`` from the two words of one more draw; a *core
    protocol* type (it needs the key, which the facade never delivers)."""

    def init_batch(self, cfg, keys):
        m = cfg.rng_mode
        k = jl.split_batched(keys, 2, m)
        px = jl.uniform_scalar_batched(k[:, 0], mode=m)
        py = jl.uniform_scalar_batched(k[:, 1], mode=m)
        k2 = jl.split_batched(k[:, 1], 2, m)
        vx = (f32(0.02) * jl.uniform_scalar_batched(k2[:, 0], mode=m) - f32(0.01)).astype(f32)
        vy = (f32(0.02) * jl.uniform_scalar_batched(k2[:, 1], mode=m) - f32(0.01)).astype(f32)
        n = keys.shape[0]
        return {"position": np.stack([px, py], 1).astype(f32), "velocity": np.stack([vx, vy], 1).astype(f32),
                "color": 0, "steps_taken": 0}

    def update_batch(self, s, model_state, cfg, keys):
        return self.step_batch(s, model_state)


# ===========================================================================
# C4-A  consumer / producer market  (tests/integration/test_integration.py)
# ===========================================================================

class Consumer:
    def __init__(self, base_income=1.0, propensity_to_consume=0.8):
        self.base_income = base_income
        self.propensity_to_consume = propensity_to_consume

    def init_batch(self, cfg, keys):                             # test_integration.py:33-41
        u = jl.uniform_scalar_batched(keys, mode=cfg.rng_mode)
        income = (f32(self.base_income) * (f32(0.8) + f32(0.4) * u)).astype(f32)
        return {"savings": 0.0, "consumption": 0.0, "utility": 0.0, "income": income}

    def update_batch(self, s, model_state, cfg, keys):           # test_integration.py:43-67
        price = f32(model_state["env"].get("price_level", 1.0))
        consumption = (f32(self.propensity_to_consume) * s["income"] / price).astype(f32)
        savings = (s["savings"] + (s["income"] - consumption * price)).astype(f32)
        utility = np.log((consumption + f32(1.0)).astype(f32)).astype(f32)
        return {"savings": savings, "consumption": consumption, "utility": utility, "income": s["income"]}


class Producer:
    def __init__(self, initial_capital=10.0, productivity=1.0, reinvestment_rate=0.3):
        self.initial_capital = initial_capital
        self.productivity = productivity
        self.reinvestment_rate = reinvestment_rate

    def init_batch(self, cfg, keys):                             # test_integration.py:85-92
        u = jl.uniform_scalar_batched(keys, mode=cfg.rng_mode)
        capital = (f32(self.initial_capital) * (f32(0.8) + f32(0.4) * u)).astype(f32)
        return {"capital": capital, "production": 0.0, "profit": 0.0}

    def update_batch(self, s, model_state, cfg, keys):           # test_integration.py:94-121
        production = (f32(self.productivity) * np.power(s["capital"], f32(0.7)).astype(f32)).astype(f32)
        price = f32(model_state["env"].get("price_level", 1.0))
        revenue = (production * price).astype(f32)
        costs = (f32(0.1) * s["capital"] + f32(0.05) * production).astype(f32)
        profit = (revenue - costs).astype(f32)
        capital = (s["capital"] + f32(self.reinvestment_rate) * profit).astype(f32)
        return {"capital": capital, "production": production, "profit": profit}


def market_update_state(env, agent_states, params, key):        # test_integration.py:125-160
    cs, ps = agent_states.get("consumers"), agent_states.get("producers")
    total_c = f32(np.sum(cs["consumption"], dtype=f32)) if cs else f32(0.0)
    total_p = f32(np.sum(ps["production"], dtype=f32)) if ps else f32(0.0)
    return market_env_from_totals(env, total_c, total_p, params)


def market_env_from_totals(env, total_c, total_p, params):
    rate = f32(params.get("price_adjustment_rate", 0.1))
    ratio = f32((total_p + f32(1e-8)) / (total_c + f32(1e-8)))
    change = f32(rate * (f32(1.0) - ratio))
    price = f32(f32(env.get("price_level", 1.0)) * (f32(1.0) + change))
    price = f32(np.clip(price, f32(0.5), f32(2.0)))
    gdp = f32(total_p * price)
    unemployment = f32(np.maximum(f32(0.0), np.minimum(f32(0.5), f32(1.0) - ratio)))
    return {"price_level": price, "gdp": gdp, "unemployment": unemployment,
            "total_consumption": total_c, "total_production": total_p}


def market_metrics(env, agent_states, params):                  # test_integration.py:163-183
    m = {"gdp": env.get("gdp", 0.0), "price_level": env.get("price_level", 1.0),
         "unemployment": env.get("unemployment", 0.0)}
    cs, ps = agent_states.get("consumers"), agent_states.get("producers")
    if cs and "utility" in cs:
        m["avg_utility"] = f32(np.mean(cs["utility"], dtype=f32))
    if ps and "profit" in ps:
        m["avg_profit"] = f32(np.mean(ps["profit"], dtype=f32))
    return m


def create_economy_model(num_consumers=20, num_producers=5, base_income=1.0,
                         propensity_to_consume=0.8, initial_capital=10.0, productivity=1.0,
                         reinvestment_rate=0.3, price_adjustment_rate=0.1, target_price=1.0,
                         seed=0, params=None, config=None):      # test_integration.py:187-283
    if params is not None:
        propensity_to_consume = params.get("propensity_to_consume", propensity_to_consume)
        productivity = params.get("productivity", productivity)
        price_adjustment_rate = params.get("price_adjustment_rate", price_adjustment_rate)
    if config is None:
        config = ModelConfig(seed=seed)
    model = Model(params={"price_adjustment_rate": price_adjustment_rate, "target_price": target_price},
                  config=config, update_state_fn=market_update_state, metrics_fn=market_metrics)
    model.add_agent_collection("consumers", AgentCollection(
        Consumer(base_income, propensity_to_consume), num_consumers))
    model.add_agent_collection("producers", AgentCollection(
        Producer(initial_capital, productivity, reinvestment_rate), num_producers))
    for k, v in {"price_level": 1.0, "gdp": 0.0, "unemployment": 0.0,
                 "total_consumption": 0.0, "total_production": 0.0}.items():
        model.add_env_state(k, v)
    return model


# ===========================================================================
# C5  sweep model  (tests/unit/test_analysis.py)
# ===========================================================================

class GrowthAgent:
    """``DummyAgent`` of ``test_analysis.py:22-39``: ``value *= 1 + growth_rate``."""

    def __init__(self, growth_rate=0.1, initial_value=0.0):
        self.growth_rate = growth_rate
        self.initial_value = initial_value

    def init_batch(self, cfg, keys):
        return {"value": unbatched(f32(self.initial_value))}

    def update_batch(self, s, model_state, cfg, keys):
        return {"value": (s["value"] * f32(1.0 + self.growth_rate)).astype(f32)}


def create_test_model(growth_rate=0.1, adjustment_rate=0.1, initial_value=0.0, num_agents=10,
                      seed=0, params=None, config=None):        # test_analysis.py:43-144
    if params is not None:
        growth_rate = params.get("growth_rate", growth_rate)
        adjustment_rate = params.get("adjustment_rate", adjustment_rate)
    if config is None:
        config = ModelConfig(seed=seed)
    mparams = {"adjustment_rate": adjustment_rate, "target_price": 1.2}

    def update_fn(env, agent_states, params, key):
        # test_analysis.py:105-111: env values and params are Python floats here, so the
        # recursion is float64 arithmetic (nothing touches a jnp array)
        new = dict(env)
        new["price_level"] += params["adjustment_rate"] * (params["target_price"] - env["price_level"])
        return new

    def metrics_fn(env, agent_states, params):
        m = {}
        cs = agent_states.get("consumers")
        if cs and "value" in cs:
            m["avg_value"] = f32(np.mean(cs["value"], dtype=f32))
        m["price_level"] = env["price_level"]
        # jnp.abs(python_float - python_float): the difference is formed in float64, then
        # converted to float32 by jnp.abs
        m["price_gap"] = f32(np.abs(f32(env["price_level"] - params["target_price"])))
        return m

    model = Model(params=mparams, config=config, update_state_fn=update_fn, metrics_fn=metrics_fn)
    model.add_agent_collection("consumers", AgentCollection(GrowthAgent(growth_rate, initial_value), num_agents))
    model.add_env_state("price_level", 1.0)
    model.add_env_state("interest_rate", 0.05)
    return model


# ===========================================================================
# contract models from the reference's unit tests
# ===========================================================================

class WealthAgent:
    """``TestAgent`` of ``tests/unit/test_agent.py:44-100``."""

    def init_batch(self, cfg, keys):
        m = cfg.rng_mode
        k = jl.split_batched(keys, 2, m)
        return {"wealth": jl.uniform_scalar_batched(k[:, 0], 0.0, 100.0, m),
                "productivity": jl.uniform_scalar_batched(k[:, 1], 0.5, 1.5, m)}

    def update_batch(self, s, model_state, cfg, keys):
        income = (s["productivity"] * f32(model_state.get("wage_rate", 1.0))).astype(f32)
        return {"wealth": (s["wealth"] + income).astype(f32), "productivity": s["productivity"]}


class IncrementAgent:
    """``DummyAgent`` of ``tests/unit/test_model.py:43-48``."""

    def init_batch(self, cfg, keys):
        return {"value": jl.uniform_scalar_batched(keys, 0.0, 10.0, cfg.rng_mode)}

    def update_batch(self, s, model_state, cfg, keys):
        return {"value": (s["value"] + f32(model_state["env"].get("increment", 1.0))).astype(f32)}


def counter_update_fn(env, agent_states, params, key):          # test_model.py:20-27 / conftest.py:78-83
    new = dict(env)
    new["counter"] = env.get("counter", 0) + 1
    return new


def counter_metrics_fn(env, agent_states, params):              # test_model.py:28-40
    m = {}
    cs = agent_states.get("consumers")
    if cs and "value" in cs:
        m["total_value"] = f32(np.sum(cs["value"], dtype=f32))
    m["step_counter"] = env.get("counter", 0)
    return m


class SimpleAgent:
    """``tests/conftest.py:55-75``."""
    growth_rate = 0.1

    def init_batch(self, cfg, keys):
        n = keys.shape[0]
        ids = np.array([jl.randint(keys[i], (), 0, 1000000, cfg.rng_mode) for i in range(n)], dtype=i32)
        return {"value": unbatched(f32(0.0)), "id": ids}

    def update_batch(self, s, model_state, cfg, keys):
        return {"value": (s["value"] * f32(1.0 + self.growth_rate)).astype(f32), "id": s["id"]}


# ===========================================================================
# C2  Schelling segregation (builder-authored rule; DESIGN.md "Schelling rule")
# ===========================================================================

def moore_counts(grid, periodic):
    """For every cell: (#occupied, #type0, #type1) among its 8 Moore neighbours."""
    occ = (grid >= 0).astype(i32)
    t0 = (grid == 0).astype(i32)
    t1 = (grid == 1).astype(i32)

    def nsum(a):
        if periodic:
            tot = np.zeros_like(a)
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    if dx or dy:
                        tot += np.roll(np.roll(a, dx, axis=0), dy, axis=1)
            return tot
        p = np.pad(a, 1)
        H, W = a.shape
        tot = np.zeros_like(a)
        for dx in (0, 1, 2):
            for dy in (0, 1, 2):
                if dx != 1 or dy != 1:
                    tot += p[dx:dx + H, dy:dy + W]
        return tot

    return nsum(occ), nsum(t0), nsum(t1)


class SchellingAgent:
    """State layout of ``SchellingSocialAgent`` (``schelling_model.py:26-31``):
    ``type i32, position i32[2], satisfied bool, moves i32``.  The rule is collective
    (a matching of movers to empty cells), so it is expressed on whole columns with
    the collection key -- it is not a per-agent ``vmap`` body.

    DESIGN.md "Schelling rule": (1) every agent is satisfied iff it has no occupied Moore
    neighbour or same/occupied >= threshold (float32), on the pre-step grid; (2) U = the
    unsatisfied agents by ascending cell id, E = env['empty_cells'] in slot order; (3) with
    rk = bits(coll_key, (8,)), mover k < min(|U|, |E|) is the agent U[piU(k)], it moves to
    the cell in slot piE(k) and its old cell takes that slot."""

    def __init__(self, similarity_threshold=0.5):
        self.similarity_threshold = similarity_threshold
        self.last_moves = None       # (slots, old cells) of the latest step, for the env phase

    def init_batch(self, cfg, keys):                             # schelling_model.py:26-31
        return {"type": 0, "position": unbatched(np.zeros(2, dtype=i32)),
                "satisfied": False, "moves": 0}

    def update_collective(self, s, model_state, cfg, coll_key):
        env = model_state["env"]
        grid = np.asarray(env["grid"], dtype=i32)
        periodic = bool(env.get("grid_periodic", False))
        G0, G1 = grid.shape
        occ, t0, t1 = moore_counts(grid, periodic)
        x, y = s["position"][:, 0], s["position"][:, 1]
        a_occ = occ[x, y]
        a_same = np.where(s["type"] == 0, t0[x, y], t1[x, y])
        with np.errstate(divide="ignore", invalid="ignore"):
            frac = (a_same.astype(f32) / a_occ.astype(f32)).astype(f32)
        satisfied = (a_occ == 0) | (frac >= f32(self.similarity_threshold))
        cell = x.astype(np.int64) * G1 + y
        unsat = np.nonzero(~satisfied)[0]
        U = unsat[np.argsort(cell[unsat], kind="stable")]        # unsatisfied agents, by cell id
        ec = np.asarray(env["empty_cells"], dtype=np.int64).reshape(-1, 2)
        E = ec[:, 0] * G1 + ec[:, 1]                             # empty cells, slot order
        u, e = len(U), len(E)
        m = min(u, e)
        rk = jl.random_bits(coll_key, (8,), cfg.rng_mode)
        k = np.arange(m, dtype=np.uint32)
        src = U[jl.feistel_permute(k, u, rk[0:4]).astype(np.int64)]
        slots = jl.feistel_permute(k, e, rk[4:8]).astype(np.int64)
        dst = E[slots]
        self.last_moves = (slots, cell[src])
        pos = s["position"].copy()
        pos[src, 0] = (dst // G1).astype(i32)
        pos[src, 1] = (dst % G1).astype(i32)
        moves = s["moves"].copy()
        moves[src] += 1
        return {"type": s["type"], "position": pos, "satisfied": satisfied, "moves": moves}


def make_schelling_update_state(agent_type):
    def schelling_update_state(env, agent_states, params, key):
        """Env phase: segregation index from the pre-step grid, grid rebuilt from the moved agents,
        vacated cells written into the empty-cell slots their movers took
        (``schelling_model.py:119-139`` layout)."""
        a = agent_states["agents"]
        grid = np.asarray(env["grid"], dtype=i32)
        G1 = grid.shape[1]
        occ, t0, t1 = moore_counts(grid, bool(env.get("grid_periodic", False)))
        same = np.where(grid == 0, t0, t1)
        sel = (grid >= 0) & (occ > 0)
        ratios = (same[sel].astype(f32) / occ[sel].astype(f32)).astype(f32)
        seg = f32(np.sum(ratios, dtype=np.float64) / max(1, ratios.size))
        new_grid = -np.ones_like(grid)
        new_grid[a["position"][:, 0], a["position"][:, 1]] = a["type"]
        ec = np.asarray(env["empty_cells"], dtype=i32).reshape(-1, 2).copy()
        slots, old_cells = agent_type.last_moves
        ec[slots, 0] = (old_cells // G1).astype(i32)
        ec[slots, 1] = (old_cells % G1).astype(i32)
        new = dict(env)
        new["grid"] = new_grid
        new["empty_cells"] = ec
        new["segregation_index"] = seg
        new["percent_satisfied"] = f32(np.mean(a["satisfied"], dtype=np.float64))
        new["total_moves"] = i32(np.sum(a["moves"], dtype=np.int64))
        return new
    return schelling_update_state


def schelling_metrics(env, agent_states, params):               # schelling_model.py:172-196 (keys)
    return {"percent_satisfied": env["percent_satisfied"],
            "segregation_index": env["segregation_index"],
            "total_moves": env["total_moves"]}


def schelling_initial_layout(grid_size, n_agents, ratio, seed):
    """Host-side placement: first N entries of a seeded shuffle of cell ids, types by
    ``ratio`` (``schelling_model.py:86-95,141-170``; 1-D ``permutation`` of cell ids
    instead of shuffling the (x,y) table)."""
    rng = np.random.RandomState(seed)
    cells = rng.permutation(grid_size * grid_size)[:n_agents]
    pos = np.stack([cells // grid_size, cells % grid_size], axis=1).astype(i32)
    n0 = int(n_agents * ratio)
    types = np.concatenate([np.zeros(n0, dtype=i32), np.ones(n_agents - n0, dtype=i32)])
    return types, pos


def create_schelling_model(grid_size=20, n_agents=300, ratio=0.5, similarity_threshold=0.5,
                           periodic=False, seed=42, config=None):
    if config is None:
        config = ModelConfig(seed=seed)
    types, pos = schelling_initial_layout(grid_size, n_agents, ratio, seed)
    agent_type = SchellingAgent(similarity_threshold)
    coll = AgentCollection(agent_type, n_agents)
    model = Model(params={"similarity_threshold": similarity_threshold}, config=config,
                  update_state_fn=make_schelling_update_state(agent_type), metrics_fn=schelling_metrics)
    model.add_agent_collection("agents", coll)
    grid = -np.ones((grid_size, grid_size), dtype=i32)
    grid[pos[:, 0], pos[:, 1]] = types
    model.add_env_state("grid_shape", (grid_size, grid_size))
    model.add_env_state("grid_periodic", periodic)
    model.add_env_state("grid", grid)
    model.add_env_state("empty_cells", np.column_stack(np.nonzero(grid < 0)).astype(i32))
    model.add_env_state("segregation_index", 0.0)
    model.add_env_state("percent_satisfied", 0.0)
    model.add_env_state("total_moves", 0)
    model.initialize()
    # the example overwrites the broadcast defaults after init (schelling_model.py:99-115)
    coll._states["type"] = types
    coll._states["position"] = pos
    return model


# ===========================================================================
# C3  SIR on a network (builder-authored rule; DESIGN.md "SIR rule")
# ===========================================================================

SIR_KCAP = 4095


def sir_escape_table(beta):
    """q[k] = fl32(q[k-1] * fl32(1-beta)), q[0]=1: the float32 value used for (1-beta)^k."""
    q = np.empty(SIR_KCAP + 1, dtype=f32)
    q[0] = f32(1.0)
    b = f32(f32(1.0) - f32(beta))
    for k in range(1, SIR_KCAP + 1):
        q[k] = f32(q[k - 1] * b)
    return q


def edges_to_csr(n, edges):
    """CSR by source of the env edge list (``agentpy.py:557``: int32[E,2], undirected
    graphs already hold both directions, ``:581-582``)."""
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    order = np.argsort(edges[:, 0], kind="stable")
    col = edges[order, 1].astype(i32)
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    row_ptr[1:] = np.bincount(edges[:, 0], minlength=n)[:n]
    return np.cumsum(row_ptr).astype(np.int64), col


class SIRAgent:
    """``state i32`` in {0:S, 1:I, 2:R}.  Synchronous on the pre-step snapshot
    (``model.py:160``): S->I if u < 1-(1-beta)^k, I->R if u < gamma, with
    ``u = uniform(split(coll_key, N)[i])`` exactly as ``agent.py:156`` delivers keys."""

    def __init__(self, beta=0.05, gamma=0.1, initial_infected=0.01, name="agents"):
        self.beta, self.gamma, self.initial_infected, self.name = beta, gamma, initial_infected, name
        self._q = sir_escape_table(beta)

    def init_batch(self, cfg, keys):
        u = jl.uniform_scalar_batched(keys, mode=cfg.rng_mode)
        return {"state": (u < f32(self.initial_infected)).astype(i32)}

    def update_batch(self, s, model_state, cfg, keys):
        env = model_state["env"]
        row_ptr, col = env["_csr"]
        st = model_state[f"agents_{self.name}"]["state"]         # pre-step snapshot
        inf = (st == 1).astype(np.int64)
        csum = np.concatenate([[0], np.cumsum(inf[col])])
        k = (csum[row_ptr[1:]] - csum[row_ptr[:-1]]).astype(np.int64)
        u = jl.uniform_scalar_batched(keys, mode=cfg.rng_mode)
        p_inf = (f32(1.0) - self._q[np.minimum(k, SIR_KCAP)]).astype(f32)
        new = st.copy()
        new[(st == 0) & (u < p_inf)] = 1
        new[(st == 1) & (u < f32(self.gamma))] = 2
        return {"state": new.astype(i32)}


def sir_metrics(env, agent_states, params):
    s = next(iter(agent_states.values()))["state"]
    return {"count_S": i32(np.sum(s == 0)), "count_I": i32(np.sum(s == 1)), "count_R": i32(np.sum(s == 2))}


def create_sir_model(n, edges, beta=0.05, gamma=0.1, initial_infected=0.01, seed=42, config=None):
    if config is None:
        config = ModelConfig(seed=seed)
    model = Model(params={"beta": beta, "gamma": gamma}, config=config, metrics_fn=sir_metrics)
    model.add_agent_collection("agents", AgentCollection(SIRAgent(beta, gamma, initial_infected), n))
    model.add_env_state("network_directed", True)
    model.add_env_state("network_edges", np.asarray(edges, dtype=i32))
    model.add_env_state("_csr", edges_to_csr(n, edges))
    return model
