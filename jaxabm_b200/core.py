"""``ModelConfig`` -- same four fields and defaults as ``jaxabm/core.py:55-86``."""
from __future__ import annotations


def has_jax() -> bool:
    """``jaxabm/core.py:30-40``.  This engine does not use JAX; report what is importable."""
    try:
        import jax  # noqa: F401
        return True
    except ImportError:
        return False


class ModelConfig:
    """Configuration for model execution (``jaxabm/core.py:55-86``).

    ``rng_mode`` is an addition: which ``jax.random`` stream layout the engine reproduces
    (``None`` -> ``_native.default_rng_mode()``).
    """

    def __init__(self, seed: int = 0, steps: int = 100, track_history: bool = True,
                 collect_interval: int = 1, rng_mode=None):
        self.seed = seed
        self.steps = steps
        self.track_history = track_history
        self.collect_interval = collect_interval
        self.rng_mode = rng_mode


def show_info():
    """``jaxabm/core.py:89-117`` equivalent for this engine."""
    from . import __version__, _native as nat
    print(f"jaxabm_b200 v{__version__}")
    print("B200-native execution engine for the JaxABM agent-update hot path")
    print(f"libjxb: {nat.LIB_PATH} (ABI {nat.lib().jxb_version()})")
    print("Available components:")
    for n in ("Model", "AgentCollection", "AgentType", "SensitivityAnalysis", "ModelCalibrator"):
        print(f"  - {n}")
