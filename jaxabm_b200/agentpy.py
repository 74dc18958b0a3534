"""AgentPy-style facade: ``Agent``, ``AgentList``, ``Environment``, ``Grid``, ``Network``,
``Model``, ``Results`` (+ the sampling / analysis wrappers) with the signatures of
``jaxabm/agentpy.py:40-1485``.  It is the drop-in boundary of SURVEY.md section 8(b): it
builds a core :class:`jaxabm_b200.model.Model` inside ``run()`` exactly as
``agentpy.py:1040-1114`` does and keeps the reference's observable quirks (Appendix B).

Differences that are deliberate and documented in DESIGN.md:
* agent classes must be registered rules (``jxb_rule``); model classes name their device
  program with ``jxb_program`` -- arbitrary Python ``step`` bodies cannot be traced;
* ``add_agents`` does not build N Python ``Agent`` objects (``agentpy.py:968-974`` is an
  O(N) host loop -- 13 M objects at C2); instances are created lazily on iteration /
  ``get_agent``;
* ``Results`` additionally supports ``in`` and ``[]`` (superset; the reference's
  ``_evaluate_params_robust`` needs them, Appendix B).
"""
from __future__ import annotations

import itertools
import pickle
import time
from typing import Any, Callable, Dict, List, Optional, Tuple, Type, Union

import numpy as np

from . import random as jrandom
from .agent import AgentCollection, AgentType, UnregisteredRuleError
from .core import ModelConfig
from .model import Model as JaxModel
from .utils import convert_to_numpy, format_time

__all__ = ["Agent", "AgentList", "Environment", "Grid", "Network", "Model", "Results",
           "Parameter", "Sample", "SensitivityAnalyzer", "ModelCalibrator"]


class Agent:
    """Base class for agents (``agentpy.py:40-154``)."""

    jxb_rule: Optional[str] = None

    @classmethod
    def jxb_params(cls, p: Dict[str, Any]) -> List[float]:
        return []

    def __init__(self):
        object.__setattr__(self, "id", None)
        object.__setattr__(self, "model", None)
        object.__setattr__(self, "p", {})
        object.__setattr__(self, "_state", {})

    def setup(self) -> Dict[str, Any]:
        return {}

    def step(self, model_state: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        return self._state

    def update_state(self, new_state: Dict[str, Any]) -> None:           # agentpy.py:100-115
        if self.model is not None and hasattr(self.model, "_update_agent_state"):
            self.model._update_agent_state(self, new_state)
        else:
            self._state.update(new_state)

    def __getattr__(self, name: str) -> Any:                              # agentpy.py:117-135
        st = object.__getattribute__(self, "_state")
        if st and name in st:
            return st[name]
        raise AttributeError(f"'{self.__class__.__name__}' object has no attribute '{name}'")

    def __setattr__(self, name: str, value: Any) -> None:                 # agentpy.py:137-154
        if name in ("id", "model", "p", "_state"):
            object.__setattr__(self, name, value)
        else:
            self.update_state({name: value})


class AgentWrapper(AgentType):
    """Adapter from an :class:`Agent` class to the ``AgentType`` protocol (``agentpy.py:157-227``).
    ``setup()`` supplies the broadcast initial state; the per-agent key is ignored (F10)."""

    def __init__(self, agent_class: Type[Agent], params: Optional[Dict[str, Any]] = None):
        self.agent_class = agent_class
        self.params = params or {}
        self.agent_instance = agent_class()
        if params:
            self.agent_instance.p = params
        self.jxb_rule = getattr(agent_class, "jxb_rule", None)

    def jxb_params(self):
        return list(self.agent_class.jxb_params(self.params))

    def jxb_host_init(self, model_config) -> Dict[str, Any]:
        state = self.agent_instance.setup()
        if not isinstance(state, dict):
            if state is None:
                return {}
            raise ValueError(f"Agent.setup() must return a dictionary, got {type(state)}")
        return state

    # Plain-Python agent classes (no registered kernel): the reference's adapter verbatim (agentpy.py:183-227).
    # The rule tracer calls these ONCE on symbolic values and turns the recorded expressions into the fused
    # update kernel (jaxabm_b200/trace.py).
    def init_state(self, model_config, key):
        if self.jxb_rule is not None:
            return super().init_state(model_config, key)
        return self.jxb_host_init(model_config)

    def update(self, state, model_state, model_config, key):
        if self.jxb_rule is not None:
            return super().update(state, model_state, model_config, key)
        self.agent_instance._state = state
        new_state = self.agent_instance.step(model_state)
        if not isinstance(new_state, dict):
            if new_state is None:
                return state
            raise ValueError(f"Agent.step() must return a dictionary, got {type(new_state)}")
        return new_state


_MISSING = object()


def _scalar(v):
    """float value of a Python / NumPy scalar env entry, None for anything else."""
    if isinstance(v, (bool, int, float, np.integer, np.floating, np.bool_)):
        return float(v)
    if isinstance(v, np.ndarray) and v.ndim == 0:
        return float(v)
    return None


class _StepProbe:
    """Stands in for ``Model._jax_model`` while the user's ``step()`` is traced: reads see the host-side state,
    ``add_env_state`` calls are recorded instead of acting on the model under construction."""

    def __init__(self, real):
        self._real = real
        self.state = {"env": dict((real.state or {}).get("env", {}))}
        self.calls: List[Tuple[str, Any]] = []

    def add_env_state(self, name: str, value: Any) -> None:
        self.calls.append((name, value))
        self.state["env"][name] = value

    def __getattr__(self, name):
        return getattr(object.__getattribute__(self, "_real"), name)


class AgentList:
    """Container for one agent collection (``agentpy.py:230-378``)."""

    def __init__(self, model: "Model", n: int, agent_class: Type[Agent], **kwargs):
        self.model = model
        self.n = n
        self.agent_class = agent_class
        self.params = kwargs
        self.agent_type = AgentWrapper(agent_class, kwargs)
        self.collection = AgentCollection(agent_type=self.agent_type, num_agents=n)
        self.name: Optional[str] = None

    @property
    def states(self):
        st = self.collection.states
        return st if st is not None else {}

    def __getattr__(self, name: str) -> Any:
        if name.startswith("_") or name in ("model", "n", "agent_class", "params", "agent_type",
                                            "collection", "name"):
            raise AttributeError(name)
        st = self.states
        if name in st:
            return st[name]
        raise AttributeError(f"'AgentList' object has no attribute '{name}'")

    def __len__(self) -> int:
        return self.n

    def select(self, condition: Callable[[Any], Any]) -> "AgentList":     # agentpy.py:316-345
        st = self.states
        cols = {k: st[k] for k in st}
        mask = np.asarray(condition(cols))
        out = AgentList(self.model, int(np.sum(mask)), self.agent_class, **self.params)
        out.collection = self.collection.filter(lambda s: mask)
        return out

    def _make(self, i: int, cols=None) -> Agent:
        a = self.agent_class()
        a.model, a.id, a.p = None, i, self.params
        if cols is not None:
            a._state = {k: v[i] for k, v in cols.items()}
        a.model = None
        return a

    def __iter__(self):                                                   # agentpy.py:347-378 (lazy)
        st = self.states
        cols = {k: st[k] for k in st} if st else None
        for i in range(self.n):
            yield self._make(i, cols)


class Environment:
    """Environment state container (``agentpy.py:381-462``)."""

    def __init__(self, model: "Model"):
        object.__setattr__(self, "model", model)
        object.__setattr__(self, "state", {})

    def add_state(self, name: str, value: Any) -> None:
        self.state[name] = value
        jm = getattr(self.model, "_jax_model", None)
        if jm:
            jm.add_env_state(name, value)

    def __getattr__(self, name: str) -> Any:
        st = object.__getattribute__(self, "state")
        if name in st:
            return st[name]
        jm = getattr(object.__getattribute__(self, "model"), "_jax_model", None)
        if jm and jm.state and name in jm.state.get("env", {}):
            return jm.state["env"][name]
        raise AttributeError(f"'Environment' object has no attribute '{name}'")

    def __setattr__(self, name: str, value: Any) -> None:
        if name in ("model", "state"):
            object.__setattr__(self, name, value)
        else:
            self.add_state(name, value)


class Grid:
    """2-D grid (``agentpy.py:465-527``): shape/periodic flags in env + random placement.
    Occupancy and neighbour queries are the engine's cell binning (``csrc/schelling.cuh``)."""

    def __init__(self, model: "Model", shape: Tuple[int, int], periodic: bool = False):
        self.model = model
        self.shape = shape
        self.periodic = periodic
        model.env.add_state("grid_shape", shape)
        model.env.add_state("grid_periodic", periodic)

    def position_agents(self, agents: AgentList, positions=None) -> None:
        n = len(agents)
        width, height = self.shape
        if positions is None:                                             # agentpy.py:509-513
            key = jrandom.PRNGKey(self.model.p.get("seed", 0))
            x = jrandom.randint(key, (n,), 0, width)
            key, subkey = jrandom.split(key)
            y = jrandom.randint(subkey, (n,), 0, height)
            positions = np.column_stack((x, y))
        st = agents.collection.states
        if st is not None:                                                # agentpy.py:516 (no-op before init)
            st["position"] = np.asarray(positions, dtype=np.int32)


class Network:
    """Edge-list network in env (``agentpy.py:530-615``).  The engine bins ``network_edges`` by
    source into CSR at ``initialize()`` (``csrc/sir.cuh``)."""

    def __init__(self, model: "Model", directed: bool = False):
        self.model = model
        self.directed = directed
        model.env.add_state("network_directed", directed)
        model.env.add_state("network_edges", np.zeros((0, 2), dtype=np.int32))

    def add_edge(self, from_agent: Union[Agent, int], to_agent: Union[Agent, int]) -> None:
        a = from_agent.id if isinstance(from_agent, Agent) else from_agent
        b = to_agent.id if isinstance(to_agent, Agent) else to_agent
        cur = self.model.env.network_edges
        self.model.env.add_state("network_edges",
                                 np.concatenate([cur, np.array([[a, b]], dtype=np.int32)], axis=0))
        if not self.directed and a != b:                                  # agentpy.py:581-582
            cur = self.model.env.network_edges
            self.model.env.add_state("network_edges",
                                     np.concatenate([cur, np.array([[b, a]], dtype=np.int32)], axis=0))

    def add_edges(self, edges) -> None:
        """Bulk form of :meth:`add_edge` (same stored layout; one concatenate instead of E)."""
        e = np.asarray(edges, dtype=np.int32).reshape(-1, 2)
        if not self.directed:
            keep = e[:, 0] != e[:, 1]
            e = np.concatenate([e, e[keep][:, ::-1]], axis=0)
        self.model.env.add_state("network_edges", np.concatenate([self.model.env.network_edges, e], axis=0))

    def get_neighbors(self, agent: Union[Agent, int]) -> np.ndarray:      # agentpy.py:584-615
        i = agent.id if isinstance(agent, Agent) else agent
        e = self.model.env.network_edges
        if self.directed:
            return e[e[:, 0] == i, 1]
        return np.unique(np.concatenate([e[e[:, 0] == i, 1], e[e[:, 1] == i, 0]], axis=0))


class Results:
    """Simulation results (``agentpy.py:618-806``)."""

    class VariableContainer:
        def __init__(self, data: Dict[str, Any]):
            self._data = data
            for agent_type in {k.split(".")[1] for k in data if k.startswith("agents.")}:
                setattr(self, agent_type, self.AgentContainer(data, agent_type))

        class AgentContainer:
            def __init__(self, data: Dict[str, Any], agent_type: str):
                self._data, self._agent_type, self._variables = data, agent_type, set()
                for key in data:
                    if key.startswith(f"agents.{agent_type}."):
                        var = key.split(".")[-1]
                        self._variables.add(var)
                        setattr(self, var, self.VariableSeries(data[key], var))

            class VariableSeries:
                def __init__(self, values, name: str):
                    self._values, self._name = values, name

                def plot(self, ax=None, **kwargs):
                    import matplotlib.pyplot as plt
                    if ax is None:
                        _, ax = plt.subplots()
                    v = self._values
                    if len(v) and hasattr(v[0], "shape") and len(v[0].shape) > 0:
                        ax.plot(np.mean(np.array(v), axis=1), **kwargs)
                        ax.set_ylabel(f"Mean {self._name}")
                    else:
                        ax.plot(v, **kwargs)
                        ax.set_ylabel(self._name)
                    ax.set_xlabel("Time")
                    return ax

                def __getitem__(self, key):
                    return self._values[key]

                def __len__(self):
                    return len(self._values)

    def __init__(self, data: Dict[str, Any]):
        self._data = convert_to_numpy(data)
        self.variables = self.VariableContainer(self._data)

    def __contains__(self, key) -> bool:       # superset of the reference (Appendix B)
        return key in self._data

    def __getitem__(self, key):
        return self._data[key]

    def keys(self):
        return self._data.keys()

    def plot(self, variables: Optional[List[str]] = None, ax=None, **kwargs):
        import matplotlib.pyplot as plt
        if ax is None:
            _, ax = plt.subplots()
        if variables is None:
            for key, values in self._data.items():
                if isinstance(values, list) and all(isinstance(v, (int, float, np.number)) for v in values):
                    ax.plot(values, label=key, **kwargs)
        else:
            for var in variables:
                if var in self._data:
                    ax.plot(self._data[var], label=var, **kwargs)
        ax.legend()
        ax.set_xlabel("Time")
        return ax

    def save(self, filename: str) -> None:
        with open(filename, "wb") as f:
            pickle.dump(self._data, f)

    @classmethod
    def load(cls, filename: str) -> "Results":
        with open(filename, "rb") as f:
            return cls(pickle.load(f))


class Model:
    """Base class for models (``agentpy.py:808-1152``).

    Subclasses name the registered device program that implements their ``step`` /
    ``compute_metrics`` with the class attribute ``jxb_program`` (``'none'`` for a model
    that overrides neither)."""

    jxb_program: Optional[str] = None

    def __init__(self, parameters: Optional[Dict[str, Any]] = None, seed: Optional[int] = None):
        self.p = parameters or {}
        self.seed = seed if seed is not None else self.p.get("seed", 0)   # agentpy.py:845
        self.steps = self.p.get("steps", 100)                             # agentpy.py:848
        self.env = Environment(self)
        self._recorded_data: Dict[str, list] = {}
        self._agent_lists: Dict[str, AgentList] = {}
        self._jax_model: Optional[JaxModel] = None
        self._current_env_state: Dict[str, Any] = {}
        self._current_agent_states: Dict[str, Any] = {}
        self._running = False
        self.last_device_seconds = 0.0
        self._record_agents: Dict[str, List[str]] = {}
        self._step_has_host_effects = False       # plain-Python models: what the trace of step() found (update_state)
        self._host_env_keys: set = set()

    # ---- user hooks ------------------------------------------------------------------------
    def setup(self) -> None:
        pass

    def step(self) -> None:
        pass

    def end(self) -> None:
        pass

    def after_initialize(self) -> None:
        """Hook (not in the reference): runs once the collections exist in HBM and before
        the first step -- where a model uploads hand-built initial columns."""

    def update_state(self, env_state, agent_states, model_params, key):   # agentpy.py:895-924
        """The reference's bridge: run the user's ``step()``, then overlay ``Environment.state`` onto the env
        (which resets every facade-level env entry each step -- Appendix B).  Registered device programs never
        call it; for a plain-Python model the rule tracer runs it on symbolic values.  While it is traced,
        ``self._jax_model`` is a probe, so that host-side calls made by ``step()`` (``add_env_state`` on the core
        model, ``record``) do not act at trace time; ``run()`` replays them once per simulated step afterwards."""
        from .trace import TrKey
        self._current_env_state = env_state.copy()
        self._current_agent_states = agent_states
        if isinstance(key, TrKey):
            from .trace import Tr, TrVec
            real, probe = self._jax_model, _StepProbe(self._jax_model)
            recorded = {k: list(v) for k, v in self._recorded_data.items()}
            env_before = dict(self.env.state)
            self._jax_model = probe
            try:
                self.step()
            finally:
                self._jax_model = real
                if probe.calls or {k: len(v) for k, v in self._recorded_data.items()} != {k: len(v) for k, v in recorded.items()}:
                    self._step_has_host_effects = True
                self._recorded_data = recorded
                overlay = dict(self.env.state)
                self.env.state.clear()
                self.env.state.update(env_before)          # the trace must not leave symbolic values in the host dict
            for name, value in overlay.items():
                if isinstance(value, (Tr, TrVec)):
                    continue                               # computed from traced values: part of the kernel's tail
                old = env_before.get(name, _MISSING)
                changed = old is _MISSING or (not np.array_equal(np.asarray(value), np.asarray(old)))
                if changed:
                    # Host-side state that step() advances every step (the examples' `self.env.add_state('time',
                    # self.env.time + 1)`): it cannot be fused into the kernel -- the traced constant would freeze at its
                    # first value.  Such an entry stays a plain env READ in the kernel (no overlay), and run() steps the
                    # model one device step at a time, calling step() on the host and refreshing these env slots before
                    # each one (the reference's un-jitted loop does the same work on the host every step).
                    if _scalar(value) is None or _scalar(old if old is not _MISSING else value) is None:
                        raise UnregisteredRuleError(
                            f"{type(self).__name__}.step() changes the non-scalar Environment.state[{name!r}] on the host every "
                            "step; only scalar host-side env entries can be refreshed per step")
                    self._host_env_keys.add(name)
            for name in self._host_env_keys:
                overlay.pop(name, None)                    # the kernel reads the slot the host refreshes
        else:
            self.step()
            overlay = self.env.state
        new_env_state = {**env_state}
        for name, value in overlay.items():
            new_env_state[name] = value
        return new_env_state

    def compute_metrics(self, env_state, agent_states, model_params):
        return {}

    # ---- agents ----------------------------------------------------------------------------
    def add_agents(self, n: int, agent_class: Type[Agent], name: Optional[str] = None, **kwargs) -> AgentList:
        agent_list = AgentList(self, n, agent_class, **kwargs)
        if name is None:
            name = agent_class.__name__.lower() + "s"                    # agentpy.py:960-961
        self._agent_lists[name] = agent_list
        agent_list.name = name
        return agent_list

    def get_agent(self, collection_name: str, agent_id: int) -> Optional[Agent]:
        al = self._agent_lists.get(collection_name)
        if al is None or not (0 <= agent_id < al.n):
            return None
        st = al.states
        return al._make(agent_id, {k: st[k] for k in st} if st else None)

    def _update_agent_state(self, agent: Agent, new_state: Dict[str, Any]) -> None:
        agent._state.update(new_state)

    def record_agents(self, collection_name: str, variables) -> None:
        """Extension: fill the ``agents.<collection>.<variable>`` entries of ``Results`` that
        ``agentpy.py:1103-1106`` intends -- one device-side snapshot of the column per recorded step."""
        if isinstance(variables, str):
            variables = [variables]
        have = self._record_agents.setdefault(collection_name, [])
        have.extend(v for v in variables if v not in have)

    def record(self, name: str, value: Any) -> None:                      # agentpy.py:1031-1038
        self._recorded_data.setdefault(name, []).append(value)

    # ---- run -------------------------------------------------------------------------------
    def _program(self) -> Optional[str]:
        """Registered device program of this model class; None = plain Python, to be traced."""
        prog = type(self).jxb_program
        if prog is not None:
            return prog
        registered = [al.agent_type.jxb_rule is not None for al in self._agent_lists.values()]
        overridden = [n for n in ("step", "compute_metrics", "update_state")
                      if getattr(type(self), n) is not getattr(Model, n)]
        if registered and all(registered):
            if overridden:
                raise UnregisteredRuleError(
                    f"{type(self).__name__} overrides {overridden} in Python but its agents are registered device "
                    "rules: registered kernels and traced Python cannot be mixed in one model. Name the registered "
                    "program (class attribute jxb_program), or write the agents as plain jx.Agent classes too.")
            return "none"
        if any(registered):
            raise UnregisteredRuleError("registered agent rules and plain-Python agent classes cannot be mixed in one model")
        return None

    def run(self, steps: Optional[int] = None) -> Results:                # agentpy.py:1040-1114
        if steps is not None:
            self.steps = steps
        self._running = True
        config = ModelConfig(steps=self.steps, collect_interval=1, seed=self.seed,
                             rng_mode=self.p.get("rng_mode"))
        self.setup()
        program = self._program()
        self._step_has_host_effects = False
        self._host_env_keys = set()
        if program is None:
            # plain-Python model: exactly what agentpy.py:1071-1076 builds -- the bound update_state bridge and
            # compute_metrics; Model.initialize() traces them (and the agents' setup / step) into one kernel
            self._jax_model = JaxModel(params=self.p, config=config, update_state_fn=self.update_state,
                                       metrics_fn=self.compute_metrics)
        else:
            def _update(env_state, agent_states, params, key):  # pragma: no cover - device resident
                raise RuntimeError("device-resident model function")

            def _metrics(env_state, agent_states, params):  # pragma: no cover - device resident
                raise RuntimeError("device-resident model function")

            _update.jxb_program = program
            _metrics.jxb_program = program
            _metrics.__dict__["jxb_owner"] = self
            self._jax_model = JaxModel(params=self.p, config=config, update_state_fn=_update,
                                       metrics_fn=None if program == "none" else _metrics)
        self._jax_model._facade = self
        for name, agent_list in self._agent_lists.items():
            self._jax_model.add_agent_collection(name, agent_list.collection)
        for name, value in self.env.state.items():
            self._jax_model.add_env_state(name, value)
        for cname, variables in self._record_agents.items():
            self._jax_model.record_agent_series(cname, variables)
        start = time.time()
        self._jax_model.initialize()
        self.after_initialize()
        if program is None and self._host_env_keys:
            # step() carries per-step host-side state (see update_state): one device step at a time, step() on the host
            # and the refreshed env slots before each
            jm = self._jax_model

            def _before_step():
                jm._host_replay = True
                try:
                    self.step()
                finally:
                    jm._host_replay = False
                for name in self._host_env_keys:
                    jm._dev.set_env(jm._dev.env_index(name), _scalar(self.env.state[name]))
            self._current_env_state = dict(jm._env_state)
            self._current_agent_states = {n: c.states for n, c in jm.agent_collections.items()}
            jm._before_each_step = _before_step
            self._step_has_host_effects = False             # the per-step calls above ARE the replay
        results_dict = self._jax_model.run()
        self.last_device_seconds = self._jax_model.last_device_seconds
        if program is None and self._step_has_host_effects:
            # the host-side part of the user's step() (core-model add_env_state bookkeeping, record()) once per
            # simulated step, as the un-jitted reference loop would have run it
            self._current_env_state = dict(self._jax_model._env_state)
            self._current_agent_states = {n: c.states for n, c in self._jax_model.agent_collections.items()}
            self._jax_model._host_replay = True
            try:
                for _ in range(int(self.steps)):
                    self.step()
            finally:
                self._jax_model._host_replay = False
        elapsed = time.time() - start
        self.end()
        self._running = False
        results_dict.update(self._recorded_data)
        # agentpy.py:1103-1106: JaxModel.state never holds 'agents' -> no 'agents.*' keys (F11) unless the model
        # opted in with record_agents(): then they are the per-step snapshots of those columns [T, N, ...]
        results_dict.update(self._jax_model.agent_series)
        results = Results(results_dict)
        print(f"Simulation executed in {format_time(elapsed)}")
        return results

    def batch_run(self, parameter_ranges: Dict[str, List[Any]], repetitions: int = 1):   # agentpy.py:1116-1152
        names = list(parameter_ranges.keys())
        out = {}
        for values in itertools.product(*parameter_ranges.values()):
            params = {**self.p}
            params.update(dict(zip(names, values)))
            out[tuple(values)] = [self.__class__(params, seed=self.seed + rep).run() for rep in range(repetitions)]
        return out


# ---------------------------------------------------------------------------------------------
# sampling / analysis wrappers (agentpy.py:1160-1485) -- host glue, kept importable
# ---------------------------------------------------------------------------------------------
class Parameter:
    """``agentpy.py:1160-1207``: a named parameter with bounds and a NumPy sampler."""

    def __init__(self, name: str, bounds: Optional[Tuple[float, float]] = None, distribution: str = "uniform",
                 value: Any = None):
        self.name, self.bounds, self.distribution, self.value = name, bounds, distribution, value

    def sample(self, n: int = 1, seed: Optional[int] = None):
        if self.bounds is None:
            return [self.value] * n
        rng = np.random.RandomState(seed)
        lo, hi = self.bounds
        if self.distribution == "uniform":
            return rng.uniform(lo, hi, n)
        if self.distribution == "normal":
            mean, std = (lo + hi) / 2, (hi - lo) / 6
            return np.clip(rng.normal(mean, std, n), lo, hi)
        if self.distribution == "log-uniform":
            return np.exp(rng.uniform(np.log(lo), np.log(hi), n))
        raise ValueError(f"Unknown distribution: {self.distribution}")


def _as_param_dict(parameters) -> Dict[str, Any]:
    """The reference passes ``List[Parameter]`` (``agentpy.py:1230,1294,1404``); a ``{name: Parameter | fixed
    value}`` dict is accepted as well."""
    if isinstance(parameters, dict):
        return dict(parameters)
    return {p.name: p for p in parameters}


class Sample:
    """``agentpy.py:1210-1260``: ``Sample(parameters, n_samples)`` draws ``n_samples`` values per parameter;
    ``sample[i]`` is the i-th parameter set, ``len(sample)`` the number of sets."""

    def __init__(self, parameters, n_samples: int = 10, seed: Optional[int] = None, n: Optional[int] = None):
        if n is not None:
            n_samples = n
        self.parameters, self.n_samples, self.n, self.seed = parameters, n_samples, n_samples, seed
        self._samples: Dict[str, Any] = {}
        for i, (name, p) in enumerate(_as_param_dict(parameters).items()):
            if isinstance(p, Parameter):
                self._samples[name] = p.sample(n_samples, None if seed is None else seed + i)
            else:
                self._samples[name] = [p] * n_samples
        self.samples: List[Dict[str, Any]] = [self[j] for j in range(n_samples)]

    def __getitem__(self, index: int) -> Dict[str, Any]:                      # agentpy.py:1247-1259
        out = {}
        for name, v in self._samples.items():
            x = v[index]
            out[name] = x.item() if hasattr(x, "item") else x
        return out

    def __iter__(self):
        return iter(self.samples)

    def __len__(self):                                                        # agentpy.py:1261-1267
        return self.n_samples


class SensitivityAnalyzer:
    """``agentpy.py:1263-1380``: adapter onto :class:`jaxabm_b200.analysis.SensitivityAnalysis`."""

    def __init__(self, model_class, parameters, n_samples: int = 10,
                 metrics: Optional[List[str]] = None, seed: int = 0):
        self.model_class, self.parameters, self.n_samples = model_class, parameters, n_samples
        self.metrics, self.seed = metrics or [], seed
        pd = _as_param_dict(parameters)
        self.fixed = {k: v for k, v in pd.items() if not isinstance(v, Parameter)}
        self.ranges = {k: v.bounds for k, v in pd.items() if isinstance(v, Parameter) and v.bounds}
        self.sample = Sample(parameters, n_samples, seed=seed)                # agentpy.py:1311
        self._sa = None

    def _factory(self, params=None, config=None):
        p = {**self.fixed, **(params or {})}
        if config is not None:
            p.setdefault("seed", config.seed)
            p.setdefault("steps", config.steps)
        return self.model_class(p)

    def run(self, verbose: bool = False):
        from .analysis import SensitivityAnalysis
        self._sa = SensitivityAnalysis(self._factory, self.ranges, self.metrics, self.n_samples, self.seed)
        return self._sa.run(verbose=verbose)

    def calculate_sensitivity(self, method: str = "sobol"):
        if self._sa is None:
            raise ValueError("Must run sensitivity analysis before calculating indices")
        if method == "sobol":
            return self._sa.sobol_indices()
        return self._sa.morris_indices()     # AttributeError, as in the reference (agentpy.py:1363-1364)


class ModelCalibrator:
    """``agentpy.py:1383-1485``: adapter onto :class:`jaxabm_b200.analysis.ModelCalibrator`."""

    def __init__(self, model_class, parameters, target_metrics: Dict[str, float],
                 metrics_weights: Optional[Dict[str, float]] = None, learning_rate: float = 0.01,
                 max_iterations: int = 20, method: str = "gradient", seed: int = 0):
        self.model_class, self.parameters, self.target_metrics = model_class, parameters, target_metrics
        self.metrics_weights, self.learning_rate = metrics_weights, learning_rate
        self.max_iterations, self.method, self.seed = max_iterations, method, seed
        self._pd = _as_param_dict(parameters)
        self.fixed = {k: v for k, v in self._pd.items() if not isinstance(v, Parameter)}

    def _factory(self, params=None, config=None):
        p = {**self.fixed, **(params or {})}
        if config is not None:
            p.setdefault("seed", config.seed)
        return self.model_class(p)

    def run(self, verbose: bool = False):
        from .analysis import ModelCalibrator as Core
        init = {k: (v.value if v.value is not None else sum(v.bounds) / 2)
                for k, v in self._pd.items() if isinstance(v, Parameter)}
        bounds = {k: v.bounds for k, v in self._pd.items() if isinstance(v, Parameter) and v.bounds}
        core = Core(self._factory, init, self.target_metrics, param_bounds=bounds,
                    metrics_weights=self.metrics_weights, learning_rate=self.learning_rate,
                    max_iterations=self.max_iterations, method=self.method, seed=self.seed)   # ValueError for 'gradient'
        return core.calibrate(verbose=verbose)
