"""Model loop: same public surface as ``jaxabm/model.py:18-285``, time loop on the device.

``Model.run(steps)`` enqueues the whole loop of ``model.py:240-241`` as CUDA graphs of
fused step kernels (agent updates -> deterministic reduction -> ``update_state_fn`` +
``metrics_fn`` tail -> history row) and reads the history back once.  The PRNG key
schedule of ``model.py:129-130,156,164,183`` is reproduced bit for bit.
"""
from __future__ import annotations

import time
from typing import Any, Callable, Dict, List, Optional

import numpy as np

from . import _native as nat
from .agent import AgentCollection, UnregisteredRuleError, rule_of
from .core import ModelConfig
from .device import DeviceModel, make_desc
from . import random as jrandom

# per program: names of the Model(params=...) entries the device tail reads, with the
# defaults the reference's functions use (None = required key, KeyError if absent)
PROGRAM_PARAMS = {
    "none": [],
    "random_walk": [],          # slot 0 is the internal 'metrics live' flag
    "market": [("price_adjustment_rate", 0.1)],          # test_integration.py:136
    "growth": [("adjustment_rate", None), ("target_price", None)],   # test_analysis.py:109
    "counter": [],
    "schelling": [("similarity_threshold", 0.5)],
    "sir": [],
    "economy": [],
}
# env entries the economy program needs to find in env_state: update_environment's final dict
# comprehension (advanced_economic_model.py:1731-1735) keeps every PRE-EXISTING entry outside its
# exclusion list, and the device program bakes that in for the layout of create_economy_model
ECONOMY_ENV_REQUIRED = ("wage_rate", "price_level", "interest_rate", "gdp", "tax_rate", "energy_price",
                        "job_market_condition", "goods_availability", "consumer_goods_demand", "consumer_goods_price",
                        "capital_goods_price", "inflation_rate", "debt_to_gdp", "avg_utility", "income_per_capita",
                        "govt_spending", "energy_supply", "capital_goods_demand", "climate_impact", "pandemic_impact")
# collections a program's functions look up by name (reference dict keys)
PROGRAM_COLLECTIONS = {
    "market": {"consumer": "consumers", "producer": "producers"},
    "growth": {"growth": "consumers"},
    "counter": {"increment": "consumers", "growth": "consumers"},
    "economy": {"household": "households", "consumer_firm": "consumer_firms"},
}


def program_of(update_state_fn, metrics_fn) -> str:
    names = set()
    for fn in (update_state_fn, metrics_fn):
        if fn is None:
            continue
        p = getattr(fn, "jxb_program", None)
        if p is None:
            raise UnregisteredRuleError(
                f"{getattr(fn, '__qualname__', fn)!r} is not a registered model function: the engine runs "
                "update_state_fn / metrics_fn as the fused tail of the step kernel and needs a registered "
                f"program ({sorted(nat.PROGRAM)}); decorate with jaxabm_b200.rules.program or use the "
                "functions shipped in jaxabm_b200.rules. There is no CPU fallback.")
        names.add(p)
    if not names:
        return "none"
    if len(names) > 1:
        raise UnregisteredRuleError(f"update_state_fn and metrics_fn belong to different programs: {sorted(names)}")
    return names.pop()


class Model:
    """Core model class (``jaxabm/model.py:18``)."""

    def __init__(self, params: Optional[Dict[str, Any]] = None, config: Optional[ModelConfig] = None,
                 update_state_fn: Optional[Callable] = None, metrics_fn: Optional[Callable] = None):
        self.config = config or ModelConfig()
        self._rng = jrandom.PRNGKey(self.config.seed)                       # model.py:46
        self._agent_collections: Dict[str, AgentCollection] = {}
        self._env_state: Dict[str, Any] = {}
        self._state: Optional[Dict[str, Any]] = None
        self._params = params or {}
        self._update_state_fn = update_state_fn
        self._metrics_fn = metrics_fn
        self._time_step = 0
        self._history: List[Dict[str, Any]] = []
        self._is_initialized = False
        self._dev: Optional[DeviceModel] = None
        self._program: Optional[str] = None
        self._shard = None            # (rank, world) when split across GPUs (sharding.shard_model)
        self._shard_group = None      # how the ranks talk on the host (sharding.DistGroup)
        self._node_range = None       # (lo, hi) agents of this rank when a Network is split by node ranges
        self._pending_network = None  # env['network_edges'] set after initialize(), not yet binned into CSR
        self._record_series: List[tuple] = []        # (collection, variable) columns snapshotted with the history rows
        self.agent_series: Dict[str, Any] = {}       # 'agents.<name>.<var>' -> [n_records, N(, w)] of the last run()
        self.last_device_seconds = 0.0

    # ---- construction ---------------------------------------------------------------------
    def add_agent_collection(self, name: str, agent_collection: AgentCollection) -> None:
        if self._is_initialized:
            raise RuntimeError("Cannot add agent collections after model is initialized")   # model.py:71-72
        self._agent_collections[name] = agent_collection

    def add_env_state(self, name: str, value: Any) -> None:                 # model.py:76-99
        if getattr(self, "_host_replay", False):
            # the facade replays the host-side part of a traced Model.step() after the device run: only the
            # host copy of the env (model.py:142-144) moves, the device env is what the traced kernel left
            if self._state is not None:
                self._state.setdefault("env", {})[name] = value
            return
        if self._is_initialized and self._state is not None:
            self._state.setdefault("env", {})[name] = value
        self._env_state[name] = value
        if self._dev is not None:
            self._push_env(name, value)

    def _push_env(self, name: str, value: Any) -> None:
        dev = self._dev
        if name == "bounds":
            b = np.asarray(value, dtype=np.float32).reshape(-1)
            for nm, v in (("bounds_lo", b[0]), ("bounds_hi", b[1])):
                s = dev.env_index(nm)
                if s is not None:
                    dev.set_env(s, v)
            return
        if name == "network_edges" and self._program == "sir":
            if self._is_initialized:
                # edits after initialize() (Network.add_edge calls add_env_state once or twice per edge,
                # agentpy.py:574-582) are batched on the host: ONE CSR rebuild before the next step
                self._pending_network = value
                return
            self._set_network(value)
            return
        s = dev.env_index(name)
        if s is not None and np.ndim(value) == 0:
            dev.set_env(s, value)

    def _set_network(self, value: Any) -> None:
        if getattr(self, "_node_range", None) is not None:
            from .sharding import local_edges
            value = local_edges(value, *self._node_range)
        self._dev.set_network(value)

    def _flush_network(self) -> None:
        pending, self._pending_network = self._pending_network, None
        if pending is not None:
            self._set_network(pending)

    def record_agent_series(self, collection: str, variables) -> None:
        """Opt in to per-agent time series: every ``run()`` snapshots the named state columns on
        the device whenever it records a history row (``t % collect_interval == 0``) and ``agent_series`` /
        the facade's ``Results`` then hold ``'agents.<collection>.<variable>'`` arrays of shape
        ``[n_records, N, ...]`` -- the keys ``jaxabm/agentpy.py:1103-1106`` means to fill.  Off by default, so the
        reference's observable behaviour (no ``agents.*`` keys) is unchanged."""
        if isinstance(variables, str):
            variables = [variables]
        for v in variables:
            if (collection, v) not in self._record_series:
                self._record_series.append((collection, v))
        if self._dev is not None:
            self._apply_record_series()

    def _apply_record_series(self) -> None:
        names = list(self._agent_collections)
        pairs = []
        for cname, var in self._record_series:
            if cname not in names:
                raise KeyError(f"record_agent_series: no agent collection {cname!r}")
            t = names.index(cname)
            pairs.append((t, self._dev.field_index(t, var)))
        self._dev.record_fields(pairs)

    def model_state(self) -> Dict[str, Any]:                                # model.py:101-116
        state = {"time_step": self._time_step, "env": self._env_state}
        for name, c in self._agent_collections.items():
            state[f"agents_{name}"] = c.states
        return state

    # ---- initialisation --------------------------------------------------------------------
    def _needs_tracing(self) -> bool:
        """True when the agent types / model functions are plain user Python (no registered kernel)."""
        regs = [getattr(c.agent_type, "jxb_rule", None) is not None for c in self._agent_collections.values()]
        fns = [getattr(f, "jxb_program", None) is not None for f in (self._update_state_fn, self._metrics_fn)
               if f is not None]
        if all(regs) and all(fns):
            return False
        if any(regs) or any(fns):
            raise UnregisteredRuleError(
                "registered kernels and traced user rules cannot be mixed in one model: either every agent type / "
                "model function is a registered one (jaxabm_b200.rules) or all of them are plain Python to be traced")
        return True

    def _initialize_traced(self) -> None:
        """User Python -> generated sm_100a step kernel (jaxabm_b200/trace.py, jit.py)."""
        from . import jit
        if self._shard:
            raise UnregisteredRuleError("traced models are not population-sharded yet")
        spec, lib, variants, src = jit.build(self)
        tm = variants[-1]
        from .device import TypeSpec
        specs = [TypeSpec("traced", t["n"]) for t in tm.types]
        desc = make_desc("traced", specs, [], rng_mode=self.config.rng_mode)
        self._dev = DeviceModel(desc, traced_spec=spec, keep=(lib, spec))
        self._program = "traced"
        self._traced_source = src
        keys = jrandom.split(self._rng, len(self._agent_collections) + 1, desc.rng_mode)
        self._dev.init(self._rng)
        self._rng = keys[0]
        for i, c in enumerate(self._agent_collections.values()):
            c._attach(self._dev, i, self.config, keys[i + 1])
        if self._record_series:
            self._apply_record_series()
        self._is_initialized = True
        self._state = {"env": self._env_state.copy()}

    def initialize(self) -> None:                                           # model.py:118-144
        if not self._agent_collections:
            raise ValueError("No agent collections added to model")
        if self._needs_tracing():
            return self._initialize_traced()
        program = program_of(self._update_state_fn, self._metrics_fn)
        specs = [c.type_spec() for c in self._agent_collections.values()]
        wanted = PROGRAM_COLLECTIONS.get(program)
        if wanted:
            for name, spec in zip(self._agent_collections, specs):
                if spec.rule in wanted and wanted[spec.rule] != name:
                    raise UnregisteredRuleError(
                        f"program {program!r} looks its {spec.rule} collection up as {wanted[spec.rule]!r} "
                        f"(as the reference functions do); it was added as {name!r}")
        mparams = []
        for key, default in PROGRAM_PARAMS[program]:
            if key in self._params:
                mparams.append(float(self._params[key]))
            elif default is None:
                raise KeyError(key)
            else:
                mparams.append(float(default))
        if program == "random_walk":
            # compute_metrics looks the walkers up under a fixed collection name
            # (examples/basic_example.py:151); under any other name it reports 0.0 distances
            want = getattr(getattr(self, "_facade", None), "jxb_metrics_collection", "walkers")
            names = list(self._agent_collections)
            mparams = [1.0 if (names and names[0] == want) else 0.0]
        if program == "economy":
            missing = [k for k in ECONOMY_ENV_REQUIRED if k not in self._env_state]
            if missing:
                raise UnregisteredRuleError(
                    "the economy program implements update_environment for the env layout of "
                    f"create_economy_model; env_state lacks {missing}")
            if self._params.get("enable_climate_module") or self._params.get("enable_pandemic_module"):
                raise UnregisteredRuleError("the climate / pandemic modules are not registered device programs")
        grid = None
        if program == "schelling":
            shape = self._env_state.get("grid_shape")
            if shape is None:
                raise ValueError("Schelling needs a Grid: env state 'grid_shape' is missing")
            grid = (int(shape[0]), int(shape[1]), bool(self._env_state.get("grid_periodic", False)))
        rank, world = self._shard or (0, 1)
        self._node_range = None
        if world > 1 and program == "sir":
            # a Network splits by node ranges at multiples of 32 (whole words of the infected bitmap)
            from .dist import shard_bounds
            spec = specs[0]
            n_all = spec.n_agents
            from .sharding import network_cuts
            cuts = network_cuts(n_all, self._env_state.get("network_edges"), world)
            lo, hi = cuts[rank], cuts[rank + 1]
            if hi <= lo:
                raise ValueError(f"a network of {n_all} agents cannot be split over {world} ranks")
            spec.global_n, spec.global_offset, spec.n_agents = n_all, lo, hi - lo
            self._node_range = (lo, hi)
        elif world > 1 and program != "schelling":   # a Grid splits by row bands; every rank keeps the agent columns
            from .dist import shard_bounds
            for spec in specs:
                lo, hi = shard_bounds(spec.n_agents, rank, world)
                if hi <= lo:
                    raise ValueError(f"collection of {spec.n_agents} agents cannot be split over {world} ranks")
                spec.global_n, spec.global_offset, spec.n_agents = spec.n_agents, lo, hi - lo
        desc = make_desc(program, specs, mparams, rng_mode=self.config.rng_mode, grid=grid,
                         world_size=world, rank=rank)
        self._dev = DeviceModel(desc)
        self._program = program
        if world > 1 and program == "schelling":
            self._dev.grid_shard_setup(self._shard_group)
        for name, value in list(self._env_state.items()):
            self._push_env(name, value)
        if self._node_range is not None:
            if "network_edges" not in self._env_state:
                raise ValueError("a sharded Network model needs env state 'network_edges' before initialize()")
            self._dev.net_shard_setup(self._shard_group)
        # keys = split(_rng, C+1); _rng = keys[0]; collection i <- keys[i+1]     (model.py:129-137)
        keys = jrandom.split(self._rng, len(self._agent_collections) + 1, desc.rng_mode)
        self._dev.init(self._rng)
        if self._node_range is not None:
            self._dev.net_shard_sync()             # every rank's bitmap copy whole before anybody steps
        self._rng = keys[0]
        for i, c in enumerate(self._agent_collections.values()):
            c._attach(self._dev, i, self.config, keys[i + 1])
            host_init = getattr(c.agent_type, "jxb_host_init", None)
            if callable(host_init):
                for fname, value in host_init(self.config).items():
                    self._dev.fill(i, self._dev.field_index(i, fname), value)
        if self._record_series:
            self._apply_record_series()
        self._is_initialized = True
        self._state = {"env": self._env_state.copy()}                       # model.py:142-144

    # ---- stepping ---------------------------------------------------------------------------
    def _metric_names(self) -> List[tuple]:
        slots = list(enumerate(self._dev.metric_slots))
        if self._program == "market":
            have = {s.rule for s in (c.type_spec() for c in self._agent_collections.values())}
            if "consumer" not in have:
                slots = [x for x in slots if x[1][0] != "avg_utility"]
            if "producer" not in have:
                slots = [x for x in slots if x[1][0] != "avg_profit"]
        if self._program == "counter":
            have = set(self._agent_collections)
            if "consumers" not in have:
                slots = [x for x in slots if x[1][0] != "total_value"]
        if self._metrics_fn is None:
            return []
        return slots

    @staticmethod
    def _cast(v: float, dt):
        if dt == np.float64:
            return float(v)
        if dt == np.int32:
            return np.int32(int(v))
        return dt(v)

    def _pull_env(self) -> None:
        if self._update_state_fn is None:
            return
        for s, (name, _) in enumerate(self._dev.env_slots):
            if name in ("bounds_lo", "bounds_hi") or name.startswith("_"):
                continue
            if self._program == "economy" and name == "total_income" and self._time_step == 0:
                continue
            if name in self._env_state or self._program not in ("random_walk", "counter"):
                self._env_state[name] = self._dev.get_env(s)

    def _advance(self, steps: int, collect_interval: int) -> List[Dict[str, Any]]:
        """Run `steps` steps on the device; return the history rows they produced."""
        self._flush_network()
        rec_steps, rec, secs = self._dev.run(steps, collect_interval)
        self.last_device_seconds = secs
        self._time_step += steps
        names = self._metric_names()
        # one vectorised cast per metric column (a list of NumPy scalars, like the reference's list of
        # jnp scalars), then the rows
        cols = []
        for k, (name, dt) in names:
            col = rec[:, k]
            if dt == np.float64:
                cols.append(col.tolist())
            elif dt == np.int32:
                cols.append(list(col.astype(np.int64).astype(np.int32)))
            else:
                cols.append(list(col.astype(dt)))
        keys = [name for _, (name, _) in names]
        ts = rec_steps.tolist()
        rows = [{"time_step": ts[r], "metrics": dict(zip(keys, vals))}
                for r, vals in enumerate(zip(*cols))] if cols else [{"time_step": t, "metrics": {}} for t in ts]
        self._pull_env()
        return rows

    def step(self) -> Dict[str, Any]:                                       # model.py:146-216
        if not self._is_initialized:
            raise RuntimeError("Model must be initialized before stepping. Call initialize() first.")
        row = self._advance(1, 1)[0]
        if self.config.track_history and self._time_step % self.config.collect_interval == 0:
            self._history.append(row)                                       # model.py:206-213
        return row["metrics"]

    def run(self, steps: Optional[int] = None) -> Dict[str, List[Any]]:     # model.py:218-262
        if not self._is_initialized:
            self.initialize()
        steps_to_run = steps if steps is not None else self.config.steps
        if self.config.track_history:
            self._history = []
        start = time.time()
        ci = self.config.collect_interval if self.config.track_history else (1 << 30)
        hook = getattr(self, "_before_each_step", None)
        if hook is None:
            rows = self._advance(int(steps_to_run), ci)
        else:
            # a traced facade model whose Model.step() advances host-side state every step (agentpy.update_state):
            # the host hook runs before each device step
            rows, secs = [], 0.0
            for _ in range(int(steps_to_run)):
                hook()
                rows.extend(self._advance(1, ci))
                secs += self.last_device_seconds
            self.last_device_seconds = secs
        if self.config.track_history:
            self._history.extend(rows)
        self.agent_series = {f"agents.{c}.{v}": self._dev.series(k) for k, (c, v) in enumerate(self._record_series)}
        elapsed = max(time.time() - start, 1e-12)
        print(f"Ran {steps_to_run} steps in {elapsed:.2f}s ({steps_to_run / elapsed:.1f} steps/sec)")
        if self.config.track_history and self._history:
            metrics: Dict[str, List[Any]] = {"step": [h["time_step"] for h in self._history]}
            for name in self._history[0]["metrics"].keys():
                metrics[name] = [h["metrics"][name] for h in self._history]
            return metrics
        return {}

    @property
    def agent_collections(self) -> Dict[str, AgentCollection]:
        return self._agent_collections

    @property
    def state(self) -> Dict[str, Any]:                                      # model.py:273-285
        if self._state is None:
            self._state = {"env": self._env_state.copy()}
        return self._state

    def jit_step(self) -> Callable:
        """``model.py:287-342``: nothing in the reference calls it; here every step already is a
        compiled kernel, so this returns the bound ``step``."""
        return self.step
