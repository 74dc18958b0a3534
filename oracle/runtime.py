"""CPU restatement of the reference's agent runtime and model loop.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows, statement for statement:

* ``jaxabm/core.py:55-86``    -> :class:`ModelConfig`
* ``jaxabm/agent.py:69-243``  -> :class:`AgentCollection`
* ``jaxabm/model.py:25-285``  -> :class:`Model`

``jax.vmap`` over per-agent keys is restated as a NumPy *batch* evaluation:
agent types implement ``init_batch(config, keys[N,2])`` / ``update_batch(states,
model_state, config, keys[N,2])`` whose bodies are the reference rule applied to
whole columns (identical arithmetic, float32).  Outputs that do not depend on
the batch are broadcast exactly as ``vmap(out_axes=0)`` does (``agent.py:125-130``).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional

import numpy as np

from . import jaxlike as jl


class ModelConfig:
    """``jaxabm/core.py:55-86``."""

    def __init__(self, seed: int = 0, steps: int = 100, track_history: bool = True,
                 collect_interval: int = 1, rng_mode: Optional[int] = None):
        self.seed = seed
        self.steps = steps
        self.track_history = track_history
        self.collect_interval = collect_interval
        # not in the reference: which JAX stream layout to restate (jaxlike.MODE)
        self.rng_mode = jl.MODE if rng_mode is None else rng_mode


class Unbatched(np.ndarray):
    """Marks an output that does not depend on the mapped axis (vmap broadcasts it)."""


def unbatched(x) -> np.ndarray:
    return jl.as_jax_default(x).view(Unbatched)


def _batchify(out: Dict[str, Any], n: int) -> Dict[str, np.ndarray]:
    """``vmap(out_axes=0)`` output rule (``agent.py:125-130``): per-agent results keep
    their leading axis; anything else (Python scalars, ``unbatched(...)`` constants)
    is broadcast to ``(n, ...)`` with JAX's x64-disabled default dtype."""
    res = {}
    for k, v in out.items():
        a = jl.as_jax_default(v)
        batched = (isinstance(v, np.ndarray) and not isinstance(v, Unbatched)
                   and a.ndim >= 1 and a.shape[0] == n)
        res[k] = np.ascontiguousarray(a) if batched else np.broadcast_to(a, (n,) + a.shape).copy()
    return res


class AgentCollection:
    """``jaxabm/agent.py:54-243``."""

    def __init__(self, agent_type, num_agents: int):
        if not isinstance(num_agents, int) or num_agents <= 0:       # agent.py:83-84
            raise ValueError("num_agents must be a positive integer")
        self.agent_type = agent_type
        self.num_agents = num_agents
        self.model_config: Optional[ModelConfig] = None
        self._key = None
        self._states: Optional[Dict[str, np.ndarray]] = None

    def init(self, key, model_config: ModelConfig) -> None:          # agent.py:92-130
        if not isinstance(model_config, ModelConfig):
            raise TypeError("model_config must be a ModelConfig instance.")
        self._key = key
        self.model_config = model_config
        agent_keys = jl.split(key, self.num_agents, model_config.rng_mode)   # agent.py:115
        init_method = getattr(self.agent_type, "init_batch", None)
        if not callable(init_method):
            raise AttributeError("Agent type must implement 'init_state'")
        self._states = _batchify(init_method(model_config, agent_keys), self.num_agents)

    def update(self, model_state, key, model_config: ModelConfig) -> None:   # agent.py:132-177
        if self._states is None:
            raise ValueError("Agent collection not initialized. Call init() first.")
        if self.model_config is None:
            raise RuntimeError("Model config not set for AgentCollection.")
        collective = getattr(self.agent_type, "update_collective", None)
        if callable(collective):
            # builder-authored extension: a rule that is a function of whole columns and the
            # collection key (Schelling's mover matching); not expressible as a vmap body.
            self._states = _batchify(
                collective(self._states, model_state, model_config, key), self.num_agents)
            return
        agent_keys = jl.split(key, self.num_agents, model_config.rng_mode)   # agent.py:156
        update_method = getattr(self.agent_type, "update_batch", None)
        if not callable(update_method):
            raise AttributeError("Agent type must implement 'update'")
        self._states = _batchify(
            update_method(self._states, model_state, model_config, agent_keys), self.num_agents)

    def get_states(self):
        return self._states

    @property
    def states(self):
        return self._states

    def aggregate(self, variable: str, fn: Callable = np.mean):      # agent.py:198-211
        if variable not in self._states:
            raise ValueError(f"Variable {variable} not found in agent states")
        return fn(self._states[variable])

    def filter(self, condition) -> "AgentCollection":                # agent.py:213-243
        mask = np.asarray(condition({k: self._states[k] for k in self._states}))
        count = int(np.sum(mask))
        out = AgentCollection(self.agent_type, count)
        out.model_config = self.model_config
        out._key = self._key
        out._states = {k: v[mask] for k, v in self._states.items()}
        return out


class Model:
    """``jaxabm/model.py:18-285`` (``jit_step`` omitted: nothing calls it)."""

    def __init__(self, params=None, config: Optional[ModelConfig] = None,
                 update_state_fn=None, metrics_fn=None):
        self.config = config or ModelConfig()
        self._rng = jl.PRNGKey(self.config.seed)                     # model.py:46
        self._agent_collections: Dict[str, AgentCollection] = {}
        self._env_state: Dict[str, Any] = {}
        self._state = None
        self._params = params or {}
        self._update_state_fn = update_state_fn
        self._metrics_fn = metrics_fn
        self._time_step = 0
        self._history: List[dict] = []
        self._is_initialized = False

    def add_agent_collection(self, name, agent_collection):          # model.py:60-74
        if self._is_initialized:
            raise RuntimeError("Cannot add agent collections after model is initialized")
        self._agent_collections[name] = agent_collection

    def add_env_state(self, name, value):                            # model.py:76-99
        if self._is_initialized and self._state is not None:
            self._state.setdefault("env", {})[name] = value
        self._env_state[name] = value

    def model_state(self):                                           # model.py:101-116
        state = {"time_step": self._time_step, "env": self._env_state}
        for name, c in self._agent_collections.items():
            state[f"agents_{name}"] = c.states
        return state

    def initialize(self):                                            # model.py:118-144
        if not self._agent_collections:
            raise ValueError("No agent collections added to model")
        keys = jl.split(self._rng, len(self._agent_collections) + 1, self.config.rng_mode)
        self._rng = keys[0]
        for i, c in enumerate(self._agent_collections.values()):
            if c.model_config is None:
                c.model_config = self.config
            c.init(keys[i + 1], self.config)
        self._is_initialized = True
        self._state = {"env": self._env_state.copy()}

    def step(self):                                                  # model.py:146-216
        if not self._is_initialized:
            raise RuntimeError("Model must be initialized before stepping. Call initialize() first.")
        m = self.config.rng_mode
        self._rng, step_key = jl.split(self._rng, 2, m)              # model.py:156
        current = self.model_state()                                 # model.py:160 (pre-step snapshot)
        for name, c in self._agent_collections.items():
            step_key, coll_key = jl.split(step_key, 2, m)            # model.py:164
            c.update(current, coll_key, self.config)
        updated = {n: c.states for n, c in self._agent_collections.items()}
        if self._update_state_fn:
            step_key, update_key = jl.split(step_key, 2, m)          # model.py:183
            self._env_state = self._update_state_fn(self._env_state, updated, self._params, update_key)
        metrics = {}
        if self._metrics_fn:
            metrics = self._metrics_fn(self._env_state, updated, self._params)
        self._time_step += 1
        if self.config.track_history and self._time_step % self.config.collect_interval == 0:
            self._history.append({"time_step": self._time_step, "metrics": metrics})
        return metrics

    def run(self, steps: Optional[int] = None):                      # model.py:218-262
        if not self._is_initialized:
            self.initialize()
        n = steps if steps is not None else self.config.steps
        if self.config.track_history:
            self._history = []
        for _ in range(n):
            self.step()
        if self.config.track_history and self._history:
            out = {"step": [h["time_step"] for h in self._history]}
            for k in self._history[0]["metrics"].keys():
                out[k] = [h["metrics"][k] for h in self._history]
            return out
        return {}

    @property
    def agent_collections(self):
        return self._agent_collections

    @property
    def state(self):
        if self._state is None:
            self._state = {"env": self._env_state.copy()}
        return self._state


def key_schedule(seed: int, n_collections: int, has_env_fn: bool, steps: int, mode: int):
    """The scalar key chain of ``model.py:129-130,156,164,183`` on its own.

    Returns ``(init_keys[C,2], coll_keys[steps,C,2], update_keys[steps,2], rng_after[2])``.
    """
    rng = jl.PRNGKey(seed)
    keys = jl.split(rng, n_collections + 1, mode)
    rng = keys[0]
    init_keys = keys[1:].copy()
    coll = np.zeros((steps, n_collections, 2), dtype=np.uint32)
    upd = np.zeros((steps, 2), dtype=np.uint32)
    for t in range(steps):
        rng, step_key = jl.split(rng, 2, mode)
        for c in range(n_collections):
            step_key, ck = jl.split(step_key, 2, mode)
            coll[t, c] = ck
        if has_env_fn:
            step_key, uk = jl.split(step_key, 2, mode)
            upd[t] = uk
    return init_keys, coll, upd, rng
