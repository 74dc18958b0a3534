"""jaxabm_b200 -- B200-native execution engine for the JaxABM agent-update hot path.

Exports the names of ``jaxabm/__init__.py:61-92``.  Host side is Python with the
reference's API; the time loop runs in ``csrc/libjxb.so`` (hand-written sm_100a kernels,
C ABI in ``include/jxb.h``).  There is no CPU fallback.
"""
__version__ = "0.1.0"

from .core import ModelConfig, has_jax, show_info
from .agent import AgentType, AgentCollection, UnregisteredRuleError
from .model import Model as JaxModel
from .analysis import SensitivityAnalysis
from .analysis import ModelCalibrator as CoreModelCalibrator
from .utils import convert_to_numpy, format_time, run_parallel_simulations
from .agentpy import (Agent, AgentList, Environment, Grid, Network, Model, Results, Parameter, Sample,
                      SensitivityAnalyzer, ModelCalibrator)
from . import random, rules

jax_available = has_jax
LegacyAgent = None     # jaxabm/legacy holds empty placeholders only
LegacyModel = None

__all__ = [
    "Agent", "AgentList", "Environment", "Grid", "Network", "Model", "Results",
    "Parameter", "Sample", "SensitivityAnalyzer",
    "AgentType", "AgentCollection", "JaxModel", "ModelConfig",
    "SensitivityAnalysis", "ModelCalibrator",
    "convert_to_numpy", "format_time", "run_parallel_simulations",
    "LegacyAgent", "LegacyModel", "has_jax", "jax_available",
]
