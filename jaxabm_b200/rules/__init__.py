"""Kernel registry: the agent types and model functions that have hand-written CUDA
kernels in ``libjxb.so``.

The reference evaluates arbitrary user Python under ``jax.vmap`` (``jaxabm/agent.py:168-177``).
This engine instead ships one fused kernel per *registered* rule; the classes here are the
host-side handles (same constructor signatures, state-dict keys and collection names as
the reference code they mirror) that select those kernels.  Unregistered types raise
``UnregisteredRuleError`` -- there is no tracing and no CPU fallback.

=====================  ===============================================================
module                 mirrors
=====================  ===============================================================
``rules.random_walk``  ``examples/basic_example.py`` (C1)
``rules.market``       ``tests/integration/test_integration.py:20-283`` (C4-A)
``rules.growth``       ``tests/unit/test_analysis.py:22-144`` (C5)
``rules.contract``     ``tests/unit/test_model.py:20-48``, ``tests/unit/test_agent.py:44-100``
``rules.economy``      ``examples/models/advanced_economic_model.py`` (C4-B)
``rules.schelling``    layout of ``examples/models/schelling_model.py`` + DESIGN.md rule (C2)
``rules.sir``          layout of ``jaxabm/agentpy.py:530-615`` + DESIGN.md rule (C3)
=====================  ===============================================================
"""
from __future__ import annotations


def program(name: str):
    """Tag a model function (``update_state_fn`` / ``metrics_fn`` / facade method) with the
    registered device program that implements it."""

    def deco(fn):
        fn.jxb_program = name
        return fn

    return deco


from . import contract, economy, growth, market, random_walk, schelling, sir  # noqa: E402,F401
