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


def run_parallel_simulations(model_factory: Callable, param_sets: List[Dict[str, Any]], num_runs: int = 1,
                             seed_offset: int = 0, seeds=None, steps=None) -> List[Dict[str, Any]]:
    """``jaxabm/utils.py:128-175``: every parameter set is run ``num_runs`` times with the seeds
    ``seed_offset + i * num_runs + j``; each results dict also carries ``'params'`` and ``'seed'``, and a run
    that raises is reported and skipped.  (The reference's loop is serial despite its name; here each
    run is one device-resident time loop.)  ``seeds`` (one per parameter set) and ``steps`` are
    extensions: explicit seeds replace the schedule, ``steps`` overrides ``config.steps``."""
    from .core import ModelConfig
    plan = []
    for i, params in enumerate(param_sets):
        if seeds is not None:
            plan.append((i, 0, params, list(seeds)[i]))
        else:
            plan.extend((i, j, params, seed_offset + i * num_runs + j) for j in range(num_runs))
    runs = 1 if seeds is not None else num_runs
    all_results = []
    for i, j, params, seed in plan:
        print(f"Running simulation {i+1}/{len(param_sets)}, run {j+1}/{runs}, seed={seed}")
        try:
            model = model_factory(params=params, config=ModelConfig(seed=seed))
            results = model.run(steps) if steps is not None else model.run()
            results["params"] = params
            results["seed"] = seed
            all_results.append(results)
        except Exception as e:                     # noqa: BLE001 - as in the reference (utils.py:173-174)
            print(f"Error in simulation {i+1}/{len(param_sets)}, run {j+1}/{runs}: {e}")
    return all_results
