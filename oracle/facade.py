"""CPU restatement of the AgentPy-style facade that wraps the model loop.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``jaxabm/agentpy.py``:

* ``:157-227``   ``AgentWrapper``  (``setup``/``step`` adapted to ``init_state``/``update``;
  the per-agent key is accepted and ignored -- SURVEY F10)
* ``:230-267``   ``AgentList``
* ``:381-462``   ``Environment``
* ``:465-527``   ``Grid``;  ``:530-615`` ``Network``
* ``:895-924``   ``Model.update_state`` (calls the user's ``step()`` and then overlays
  ``Environment.state`` on the incoming env -- the overlay is what freezes
  ``time``/``mean_x``/... in ``examples/basic_example.py``)
* ``:944-976``   ``Model.add_agents`` (auto name ``cls.__name__.lower()+'s'``)
* ``:1040-1114`` ``Model.run``

Agents here are *batch* agents: ``setup()`` returns the per-agent dict (broadcast
by the runtime), ``step_batch(state, model_state)`` is the reference ``step`` body
applied to columns.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np

from . import jaxlike as jl
from .runtime import AgentCollection, Model as CoreModel, ModelConfig, unbatched


class Agent:
    def __init__(self):
        self.id = None
        self.model = None
        self.p: Dict[str, Any] = {}

    def setup(self) -> Dict[str, Any]:
        return {}

    def step_batch(self, state, model_state):
        return state


class AgentWrapper:
    """``agentpy.py:157-227``."""

    def __init__(self, agent_class, params=None):
        self.agent_class = agent_class
        self.params = params or {}
        self.agent_instance = agent_class()
        if params:
            self.agent_instance.p = params

    def init_batch(self, model_config, keys):                    # agentpy.py:175-197 (key unused)
        state = self.agent_instance.setup()
        if not isinstance(state, dict):
            if state is None:
                return {}
            raise ValueError("Agent.setup() must return a dictionary")
        return {k: unbatched(v) for k, v in state.items()}

    def update_batch(self, state, model_state, model_config, keys):   # agentpy.py:199-227
        new_state = self.agent_instance.step_batch(state, model_state)
        if new_state is None:
            return state
        return new_state


class AgentList:
    def __init__(self, model, n, agent_class, **kwargs):
        self.model = model
        self.n = n
        self.agent_class = agent_class
        self.params = kwargs
        self.agent_type = AgentWrapper(agent_class, kwargs)
        self.collection = AgentCollection(self.agent_type, n)
        self.name = None

    def __len__(self):
        return self.n


class Environment:
    def __init__(self, model):
        self.model = model
        self.state: Dict[str, Any] = {}

    def add_state(self, name, value):                            # agentpy.py:406-417
        self.state[name] = value
        if getattr(self.model, "_jax_model", None):
            self.model._jax_model.add_env_state(name, value)


class Grid:
    """``agentpy.py:465-527``."""

    def __init__(self, model, shape, periodic=False):
        self.model, self.shape, self.periodic = model, shape, periodic
        model.env.add_state("grid_shape", shape)
        model.env.add_state("grid_periodic", periodic)

    def random_positions(self, n, mode=None):                    # agentpy.py:509-513
        w, h = self.shape
        key = jl.PRNGKey(self.model.p.get("seed", 0))
        x = jl.randint(key, (n,), 0, w, mode)
        key, sub = jl.split(key, 2, mode)
        y = jl.randint(sub, (n,), 0, h, mode)
        return np.column_stack((x, y)).astype(np.int32)


class Network:
    """``agentpy.py:530-615``: int32 edge list in env; undirected stores both directions."""

    def __init__(self, model, directed=False):
        self.model, self.directed = model, directed
        model.env.add_state("network_directed", directed)
        model.env.add_state("network_edges", np.zeros((0, 2), dtype=np.int32))

    def add_edge(self, a, b):                                    # agentpy.py:559-582
        cur = self.model.env.state["network_edges"]
        self.model.env.add_state("network_edges",
                                 np.concatenate([cur, np.array([[a, b]], dtype=np.int32)], axis=0))
        if not self.directed and a != b:
            cur = self.model.env.state["network_edges"]
            self.model.env.add_state("network_edges",
                                     np.concatenate([cur, np.array([[b, a]], dtype=np.int32)], axis=0))

    def get_neighbors(self, agent_id):                           # agentpy.py:584-615
        e = self.model.env.state["network_edges"]
        if self.directed:
            return e[e[:, 0] == agent_id, 1]
        return np.unique(np.concatenate([e[e[:, 0] == agent_id, 1], e[e[:, 1] == agent_id, 0]]))


class Model:
    """``agentpy.py:808-1114`` (only what reaches the hot path)."""

    def __init__(self, parameters=None, seed=None, rng_mode=None):
        self.p = parameters or {}
        self.seed = seed if seed is not None else self.p.get("seed", 0)   # agentpy.py:845
        self.steps = self.p.get("steps", 100)                             # agentpy.py:848
        self.env = Environment(self)
        self._recorded_data: Dict[str, list] = {}
        self._agent_lists: Dict[str, AgentList] = {}
        self._jax_model: Optional[CoreModel] = None
        self._rng_mode = rng_mode

    def setup(self):
        pass

    def step(self):
        pass

    def end(self):
        pass

    def update_state(self, env_state, agent_states, model_params, key):   # agentpy.py:895-924
        self._current_env_state = dict(env_state)
        self._current_agent_states = agent_states
        self.step()
        new_env = {**env_state}
        for name, value in self.env.state.items():                        # the overlay
            new_env[name] = value
        return new_env

    def compute_metrics(self, env_state, agent_states, model_params):
        return {}

    def add_agents(self, n, agent_class, name=None, **kwargs):            # agentpy.py:944-976
        al = AgentList(self, n, agent_class, **kwargs)
        if name is None:
            name = agent_class.__name__.lower() + "s"
        self._agent_lists[name] = al
        al.name = name
        return al

    def record(self, name, value):                                        # agentpy.py:1031-1038
        self._recorded_data.setdefault(name, []).append(value)

    def run(self, steps=None):                                            # agentpy.py:1040-1114
        if steps is not None:
            self.steps = steps
        config = ModelConfig(steps=self.steps, collect_interval=1, seed=self.seed,
                             rng_mode=self._rng_mode)
        self.setup()
        self._jax_model = CoreModel(params=self.p, config=config,
                                    update_state_fn=self.update_state,
                                    metrics_fn=self.compute_metrics)
        for name, al in self._agent_lists.items():
            self._jax_model.add_agent_collection(name, al.collection)
        for name, value in self.env.state.items():
            self._jax_model.add_env_state(name, value)
        results = self._jax_model.run()
        self.end()
        results.update(self._recorded_data)
        # agentpy.py:1103-1106: JaxModel.state never has 'agents' -> no 'agents.*' keys (F11)
        return results
