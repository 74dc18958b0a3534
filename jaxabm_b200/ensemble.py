"""Ensemble dispatch: many independent runs of one model shape as ONE device launch.

This is the loop body of ``SensitivityAnalysis.run`` (``jaxabm/analysis.py:113-157``) and of
``ModelCalibrator._evaluate_params_robust`` (``jaxabm/analysis.py:434-476``), which the
reference executes as strictly serial ``model_factory(...).run()`` calls.
Here the factory is still called once per sample (cheap host objects, nothing allocated on
the device), the resulting models are checked to be *homogeneous* -- same program, same
collections, same sizes -- and the parameters that differ between them become the swept
columns of a single ``jxb_ensemble_run``.  With ``torch.distributed`` initialised, replicas
are sharded across ranks (one GPU each) with only a final gather (``dist.py``).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native as nat
from . import dist
from .device import PROGRAM_ENV, ensemble_run, make_desc
from .model import Model, PROGRAM_PARAMS, program_of

MAX_SWEPT = 8


def _signature(m: Model):
    """(program, [(name, rule, n)], type params, model params, rng_mode) of an un-run model."""
    program = program_of(m._update_state_fn, m._metrics_fn)
    specs = [(name, c.type_spec()) for name, c in m._agent_collections.items()]
    mparams = []
    for key, default in PROGRAM_PARAMS[program]:
        if key in m._params:
            mparams.append(float(m._params[key]))
        elif default is None:
            raise KeyError(key)
        else:
            mparams.append(float(default))
    if program == "random_walk":
        names = list(m._agent_collections)
        mparams = [1.0 if names and names[0] == "walkers" else 0.0]
    shape = (program, tuple((n, s.rule, s.n_agents) for n, s in specs), m.config.rng_mode,
             tuple(sorted((k, repr(v)) for k, v in m._env_state.items() if np.ndim(v) == 0)))
    return program, specs, mparams, shape


def batchable(models: Sequence[Any]) -> bool:
    if not models or not all(isinstance(m, Model) for m in models):
        return False
    if any(m._is_initialized for m in models):
        return False
    try:
        sigs = [_signature(m) for m in models]
    except Exception:
        return False
    if sigs[0][0] in ("schelling", "sir", "economy"):
        return False
    # env must still be at the program defaults the device kernel starts from
    return all(s[3] == sigs[0][3] for s in sigs)


def plan(models: Sequence[Model]):
    """-> (desc, slots, params[R, n_swept], seeds[R], env_init[16])."""
    sigs = [_signature(m) for m in models]
    program, specs0, mp0, _ = sigs[0]
    R = len(models)
    cols: List[Tuple[int, np.ndarray]] = []
    mp = np.array([s[2] for s in sigs], dtype=np.float64).reshape(R, -1)
    for k in range(mp.shape[1]):
        if np.any(mp[:, k] != mp[0, k]):
            cols.append((k, mp[:, k]))
    for ti in range(len(specs0)):
        tp = np.array([s[1][ti][1].params for s in sigs], dtype=np.float64).reshape(R, -1)
        for k in range(tp.shape[1]):
            if np.any(tp[:, k] != tp[0, k]):
                cols.append((100 + 16 * ti + k, tp[:, k]))
    if len(cols) > MAX_SWEPT:
        raise ValueError(f"more than {MAX_SWEPT} parameters differ between the replicas")
    desc = make_desc(program, [s for _, s in specs0], mp0, rng_mode=models[0].config.rng_mode)
    slots = [c[0] for c in cols]
    params = np.stack([c[1] for c in cols], axis=1) if cols else np.zeros((R, 0))
    seeds = np.array([int(m.config.seed) & 0xFFFFFFFF for m in models], dtype=np.uint32)
    return desc, slots, params, seeds, env_init(program, models[0])


ENV_DEFAULTS = {"bounds_lo": 0.0, "bounds_hi": 1.0, "mean_x": 0.5, "mean_y": 0.5, "price_level": 1.0,
                "interest_rate": 0.05, "increment": 1.0}


def env_init(program: str, m: Model) -> np.ndarray:
    """add_env_state() values of a model in the program's slot order."""
    vals = []
    for name in PROGRAM_ENV[program]:
        if name in ("bounds_lo", "bounds_hi") and "bounds" in m._env_state:
            b = np.asarray(m._env_state["bounds"], dtype=np.float32).reshape(-1)
            vals.append(float(b[0] if name == "bounds_lo" else b[1]))
        elif name in m._env_state and np.ndim(m._env_state[name]) == 0:
            vals.append(float(m._env_state[name]))
        else:
            vals.append(ENV_DEFAULTS.get(name, 0.0))
    return np.array(vals + [0.0] * (16 - len(vals)), dtype=np.float64)


def metric_layout(models: Sequence[Model]):
    """Metric (slot, name, dtype) list of the shared program, honouring absent collections."""
    from .device import DeviceModel  # noqa: F401  (layout is static; avoid allocating a device model)
    program = program_of(models[0]._update_state_fn, models[0]._metrics_fn)
    table = {
        "none": [],
        "random_walk": [("mean_x", np.float64), ("mean_y", np.float64), ("mean_distance", np.float32),
                        ("max_distance", np.float32), ("num_red", np.int32), ("num_blue", np.int32),
                        ("time", np.int32)],
        "market": [("gdp", np.float32), ("price_level", np.float32), ("unemployment", np.float32),
                   ("avg_utility", np.float32), ("avg_profit", np.float32)],
        "growth": [("avg_value", np.float32), ("price_level", np.float64), ("price_gap", np.float32)],
        "counter": [("total_value", np.float32), ("step_counter", np.int32)],
    }[program]
    out = list(enumerate(table))
    if program == "market":
        rules = {c.type_spec().rule for c in models[0]._agent_collections.values()}
        if "consumer" not in rules:
            out = [x for x in out if x[1][0] != "avg_utility"]
        if "producer" not in rules:
            out = [x for x in out if x[1][0] != "avg_profit"]
    if models[0]._metrics_fn is None:
        out = []
    return out


def run_last_metrics(models: Sequence[Model], steps: Optional[int] = None):
    """Final value of every metric for every model -> (dict name -> array[R], device_seconds).

    Sharded over ranks when torch.distributed is initialised with world_size > 1."""
    steps = models[0].config.steps if steps is None else steps
    desc, slots, params, seeds, env0 = plan(models)
    layout = metric_layout(models)
    R = len(models)
    lo, hi = dist.shard_range(R)
    secs = 0.0
    local = np.zeros((hi - lo, nat.MAX_METRICS), dtype=np.float64)
    if hi > lo:
        vals, secs = ensemble_run(desc, slots, params[lo:hi], seeds[lo:hi], steps, env0)
        local[:, :vals.shape[1]] = vals
    full = dist.gather_rows(local, R)
    secs = dist.max_over_ranks(secs)
    out = {}
    for k, (name, dt) in layout:
        col = full[:, k]
        out[name] = col if dt == np.float64 else col.astype(dt)
    return out, secs


def run_models(models: Sequence[Any], steps: Optional[int] = None, full_history: bool = False) -> List[Dict[str, Any]]:
    """Run every model; one ensemble launch when they are batchable and only the last
    values are wanted, else one device run each.  Returns one results dict per model."""
    if not full_history and batchable(models):
        last, _ = run_last_metrics(models, steps)
        n_steps = models[0].config.steps if steps is None else steps
        return [{"step": [n_steps], **{k: [v[i]] for k, v in last.items()}} for i in range(len(models))]
    return [m.run(steps) if steps is not None else m.run() for m in models]
