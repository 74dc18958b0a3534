"""Host utilities with the semantics of ``jaxabm/utils.py`` (no acceleration value)."""
from __future__ import annotations

from typing import Any, Callable, Dict, List

import numpy as np


def convert_to_numpy(data: Any) -> Any:
    """``jaxabm/utils.py:17-38``: recursive conversion; engine values already are NumPy."""
    if isinstance(data, dict):
        return {k: convert_to_numpy(v) for k, v in data.items()}
    if isinstance(data, list):
        return [convert_to_numpy(v) for v in data]
    if isinstance(data, tuple):
        return tuple(convert_to_numpy(v) for v in data)
    if hasattr(data, "__array__") and not isinstance(data, np.ndarray) and not np.isscalar(data):
        return np.asarray(data)
    return data


def is_valid_params(params: Dict[str, Any], required_keys: List[str]) -> bool:
    return all(k in params for k in required_keys)


def format_time(seconds: float) -> str:
    """Same output format as ``jaxabm/utils.py:54-74`` ('12.34s', '2m 5.00s', '1h 2m 3.00s')."""
    if seconds < 60:
        return f"{seconds:.2f}s"
    hours, rest = divmod(seconds, 3600) if seconds >= 3600 else (0, seconds)
    minutes = int(rest / 60)
    secs = rest % 60
    if hours:
        return f"{int(hours)}h {minutes}m {secs:.2f}s"
    return f"{minutes}m {secs:.2f}s"


def mean_over_runs(results_list: List[Dict[str, Any]]) -> Dict[str, Any]:
    """``jaxabm/utils.py:76-105``: per-metric mean over runs whose series have equal length."""
    if not results_list:
        return {}
    out = {}
    first = results_list[0]
    for k in first:
        if all(k in r and len(r[k]) == len(first[k]) for r in results_list):
            out[k] = np.mean([np.array(r[k]) for r in results_list], axis=0).tolist()
    return out


def standardize_metrics(metrics: Dict[str, Any]) -> Dict[str, float]:
    """``jaxabm/utils.py:108-125``."""
    out = {}
    for k, v in metrics.items():
        if hasattr(v, "item"):
            out[k] = float(v.item())
        elif isinstance(v, (int, float)):
            out[k] = float(v)
    return out


def run_parallel_simulations(model_factory: Callable, param_sets: List[Dict[str, Any]], seeds=None,
                             steps=None) -> List[Dict[str, Any]]:
    """``jaxabm/utils.py:128-175`` -- but actually batched: homogeneous engine models are
    dispatched as one ensemble launch (``jaxabm_b200.ensemble``)."""
    from .core import ModelConfig
    from .ensemble import run_models
    seeds = list(seeds) if seeds is not None else list(range(len(param_sets)))
    models = [model_factory(params=p, config=ModelConfig(seed=s)) for p, s in zip(param_sets, seeds)]
    return run_models(models, steps=steps, full_history=True)
