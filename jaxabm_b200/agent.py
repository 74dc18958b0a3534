"""Agent runtime: ``AgentType`` protocol and ``AgentCollection`` -- same public surface as
``jaxabm/agent.py:19-243``, with the struct-of-arrays state living in HBM.

``AgentCollection.init`` / ``update`` are the reference's ``split(key, N)`` + ``vmap``
(``agent.py:115,125`` / ``:156,173``) executed as one fused CUDA kernel per collection;
per-agent keys are derived in registers.  Only *registered* agent types can run (the
reference traces arbitrary Python with JAX; this engine ships hand-written kernels, see
``jaxabm_b200.rules``) -- anything else raises :class:`UnregisteredRuleError`, never a
CPU fallback.
"""
from __future__ import annotations

from collections.abc import MutableMapping
from typing import Any, Callable, Dict, Optional

import numpy as np

from . import _native as nat
from .core import ModelConfig
from .device import DeviceModel, TypeSpec, make_desc

AgentState = Dict[str, Any]


class UnregisteredRuleError(NotImplementedError):
    """The agent type / model function has no registered CUDA kernel."""


class AgentType:
    """Protocol of ``jaxabm/agent.py:19-51``.

    Engine-backed types additionally carry ``jxb_rule`` (kernel name) and
    ``jxb_params()`` (the rule constants, in the kernel's order).
    """

    jxb_rule: Optional[str] = None

    def jxb_params(self):
        return []

    def init_state(self, model_config: Any, key: Any) -> AgentState:
        raise UnregisteredRuleError(
            f"{type(self).__name__}.init_state runs on the device for all agents at once "
            "(AgentCollection.init); it has no per-agent Python body in this engine")

    def update(self, state: AgentState, model_state: Dict[str, Any], model_config: Any, key: Any) -> AgentState:
        raise UnregisteredRuleError(
            f"{type(self).__name__}.update runs on the device for all agents at once "
            "(AgentCollection.update); it has no per-agent Python body in this engine")


def rule_of(agent_type) -> str:
    rule = getattr(agent_type, "jxb_rule", None)
    if rule is None or rule not in nat.RULE:
        name = getattr(agent_type, "__name__", type(agent_type).__name__)
        raise UnregisteredRuleError(
            f"agent type {name!r} has no registered CUDA rule (jxb_rule). Registered rules: "
            f"{sorted(nat.RULE)}. See jaxabm_b200.rules, or write the type as plain Python against "
            "jaxabm_b200.numpy / jaxabm_b200.random and add it to a Model whose functions are plain Python too: "
            "Model.initialize() then traces it into a generated kernel. There is no CPU fallback.")
    return rule


class DeviceStates(MutableMapping):
    """``AgentCollection.states``: name -> ``(N, ...)`` array view of the HBM columns.

    Reads download the column; assignments upload it (the reference mutates
    ``collection._states[...]`` directly, e.g. ``jaxabm/agentpy.py:516-527``).
    """

    def __init__(self, dev: DeviceModel, tidx: int, on_write: Optional[Callable[[str], None]] = None):
        self._dev, self._t, self._on_write = dev, tidx, on_write

    def __getitem__(self, name):
        return self._dev.download(self._t, self._dev.field_index(self._t, name))

    def __setitem__(self, name, value):
        self._dev.upload(self._t, self._dev.field_index(self._t, name), value)
        if self._on_write:
            self._on_write(name)

    def __delitem__(self, name):
        raise TypeError("agent state columns cannot be deleted")

    def __iter__(self):
        return (n for n, _, _ in self._dev.fields[self._t])

    def __len__(self):
        return len(self._dev.fields[self._t])

    def __bool__(self):
        return True


class AgentCollection:
    """Collection of agents of one type (``jaxabm/agent.py:54-243``)."""

    def __init__(self, agent_type: AgentType, num_agents: int):
        if not isinstance(num_agents, int) or isinstance(num_agents, bool) or num_agents <= 0:
            raise ValueError("num_agents must be a positive integer")          # agent.py:83-84
        self.agent_type = agent_type
        self.num_agents = num_agents
        self.model_config: Optional[ModelConfig] = None
        self._key = None
        self._dev: Optional[DeviceModel] = None
        self._tidx = 0
        self._initialized = False
        self._pending: Dict[str, Any] = {}

    # -- wiring used by Model.initialize -------------------------------------------------
    def _attach(self, dev: DeviceModel, tidx: int, config: ModelConfig, key=None) -> None:
        self._dev, self._tidx = dev, tidx
        self.model_config = config
        self._key = key
        self._initialized = True

    def type_spec(self) -> TypeSpec:
        return TypeSpec(rule_of(self.agent_type), self.num_agents, self.agent_type.jxb_params())

    # -- reference API ---------------------------------------------------------------------
    def init(self, key: Any, model_config: ModelConfig) -> None:
        """``agent.py:92-130``: per-agent keys ``split(key, N)`` and the type's ``init_state``,
        evaluated for every agent by one kernel."""
        if not isinstance(model_config, ModelConfig):
            raise TypeError("model_config must be a ModelConfig instance.")   # agent.py:103-104
        if not isinstance(self.num_agents, int) or self.num_agents <= 0:
            raise ValueError("Number of agents must be a positive integer.")
        if not callable(getattr(self.agent_type, "init_state", None)):
            raise AttributeError(f"Agent type {getattr(self.agent_type, '__name__', type(self.agent_type).__name__)} "
                                 "must implement 'init_state'")           # agent.py:118-120
        self._key = key
        self.model_config = model_config
        if self._dev is None:
            # stand-alone collection (tests/unit/test_agent.py): a one-collection device model
            desc = make_desc("none", [self.type_spec()], rng_mode=model_config.rng_mode)
            self._dev, self._tidx = DeviceModel(desc), 0
        self._dev.collection_init(self._tidx, key)
        host_init = getattr(self.agent_type, "jxb_host_init", None)
        if callable(host_init):
            for name, value in host_init(model_config).items():
                self._dev.fill(self._tidx, self._dev.field_index(self._tidx, name), value)
        self._initialized = True

    def update(self, model_state: Dict[str, Any], key: Any, model_config: ModelConfig) -> None:
        """``agent.py:132-177``: one fused update of this collection with the caller's key."""
        if not self._initialized or self._dev is None:
            raise ValueError("Agent collection not initialized. Call init() first.")   # agent.py:150-151
        if self.model_config is None:
            raise RuntimeError("Model config not set for AgentCollection. Ensure Model.initialize() was called.")
        if not callable(getattr(self.agent_type, "update", None)):
            raise AttributeError("Agent type must implement 'update'")
        bind = getattr(self.agent_type, "jxb_bind_model_state", None)
        if callable(bind):
            bind(self._dev, self._tidx, model_state or {})
        self._dev.collection_update(self._tidx, key)

    def get_states(self):
        return self.states

    @property
    def states(self) -> Optional[DeviceStates]:
        if not self._initialized or self._dev is None:
            return None
        return DeviceStates(self._dev, self._tidx)

    @property
    def _states(self):
        return self.states

    @_states.setter
    def _states(self, value):
        if value is None:
            return
        if not self._initialized or self._dev is None:
            # reference: whatever is assigned before init() is overwritten by init()
            self._pending = dict(value)
            return
        st = self.states
        for k, v in dict(value).items():
            st[k] = v

    def aggregate(self, variable: str, fn: Callable = np.mean) -> Any:       # agent.py:198-211
        st = self.states
        if st is None or variable not in st:
            raise ValueError(f"Variable {variable} not found in agent states")
        return fn(st[variable])

    def filter(self, condition: Callable[[Dict[str, Any]], Any]) -> "AgentCollection":   # agent.py:213-243
        """New collection holding the agents for which ``condition(states)`` is true, in their order
        (``{k: v[mask]}`` of the reference) -- a stream compaction on the device: flags from the traced
        condition (``jaxabm_b200/select.py``), exclusive scan, one ordered scatter per column into the new
        collection's HBM columns.  A condition the tracer cannot follow is evaluated on the host over the
        columns it reads (downloaded lazily) and only the mask travels back."""
        from . import select
        from .trace import TraceError
        if not self._initialized or self._dev is None:
            raise ValueError("Agent collection not initialized. Call init() first.")
        dev, t = self._dev, self._tidx
        try:
            count = dev.filter_select(t, program=select.compile_predicate(condition, dev.fields[t]))
        except TraceError:
            st = self.states

            class _Lazy(dict):
                def __missing__(self, k):
                    self[k] = st[k]
                    return self[k]

                def __iter__(self):
                    return iter(st)
            mask = np.asarray(condition(_Lazy()))
            if mask.dtype != np.bool_ or mask.shape != (self.num_agents,):
                raise ValueError("filter condition must return one boolean per agent")
            count = dev.filter_select(t, mask=mask)
        out = AgentCollection(self.agent_type, count)   # ValueError when nothing matches, as in the reference
        out.model_config = self.model_config
        out._key = self._key
        spec = out.type_spec() if getattr(self.agent_type, "jxb_rule", None) else None
        if spec is None:
            raise UnregisteredRuleError("filter on a traced collection is not supported yet")
        desc = make_desc("none", [spec], rng_mode=self.model_config.rng_mode if self.model_config else None)
        out._dev, out._tidx, out._initialized = DeviceModel(desc), 0, True
        dev.filter_gather(t, out._dev, 0)
        return out
