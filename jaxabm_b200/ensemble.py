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


def _traced(models: Sequence[Any]) -> bool:
    try:
        return all(isinstance(m, Model) and not m._is_initialized and m._needs_tracing() for m in models)
    except Exception:
        return False


def batchable(models: Sequence[Any]) -> bool:
    if not models or not all(isinstance(m, Model) for m in models):
        return False
    if any(m._is_initialized for m in models):
        return False
    if _traced(models):
        return True
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


def _run_traced(models: Sequence[Model], steps: int):
    """Replica-parallel run of user-written (traced) models: every model is traced (cheap), all must
    generate the same source -- they then differ only in their constant tables, env values and seeds --
    and the generated library's ensemble kernel runs them all in one launch."""
    import ctypes as C
    from . import jit, trace as T
    R = len(models)
    lo, hi = dist.shard_range(R)
    src0 = meta0 = variants0 = None
    consts, env0 = [], []
    for m in models:
        variants = jit.trace_variants(m)
        src, meta = T.generate_source(variants)
        if src0 is None:
            src0, meta0, variants0 = src, meta, variants
        elif src != src0:
            raise ValueError("the models of an ensemble must trace to the same kernel (same structure, same sizes)")
        consts.append(meta["consts"])
        tm = variants[-1]
        row = np.zeros(32, dtype=np.float64)
        for k, name in enumerate(tm.env_names):
            v = m._env_state.get(name, 0.0)
            row[k] = float(v) if T._env_dtype_of(v) is not None else 0.0
        env0.append(row)
    lib = jit.compile_source(src0)
    tm = variants0[-1]
    n_agents = (C.c_longlong * 4)(*([t["n"] for t in tm.types] + [0] * (4 - len(tm.types))))
    nc = len(meta0["consts"])
    cst = np.ascontiguousarray(np.array(consts, dtype=np.float64).reshape(R, max(nc, 0)))
    envs = np.ascontiguousarray(np.stack(env0, axis=0))
    seeds = np.array([int(m.config.seed) & 0xFFFFFFFF for m in models], dtype=np.uint32)
    local = np.zeros((hi - lo, nat.MAX_METRICS), dtype=np.float64)
    secs = C.c_double(0.0)
    if hi > lo:
        lib.jxc_ensemble_run.restype = C.c_int
        rc = lib.jxc_ensemble_run(C.c_int(nat.engine().device), C.c_int(hi - lo), C.c_int(int(steps)), n_agents,
                                  nat.ptr(np.ascontiguousarray(cst[lo:hi])), C.c_int(nc),
                                  nat.ptr(np.ascontiguousarray(envs[lo:hi])), nat.ptr(np.ascontiguousarray(seeds[lo:hi])),
                                  C.c_int(1 if tm.has_env_fn else 0), C.c_int(int(models[0].config.rng_mode
                                                                                 if models[0].config.rng_mode is not None
                                                                                 else nat.default_rng_mode())),
                                  nat.ptr(local), C.byref(secs))
        if rc != 0:
            raise nat.JxbError(rc, "traced ensemble launch failed")
    full = dist.gather_rows(local, R)
    out = {}
    for k, (name, v) in enumerate(tm.metrics):
        dt = {T.F32: np.float32, T.I32: np.int32, T.WI: np.int32, T.BOOL: np.bool_}.get(v.dtype, np.float64)
        col = full[:, k]
        out[name] = col if dt == np.float64 else col.astype(dt)
    return out, dist.max_over_ranks(secs.value)


def run_last_metrics(models: Sequence[Model], steps: Optional[int] = None):
    """Final value of every metric for every model -> (dict name -> array[R], device_seconds).

    Sharded over ranks when torch.distributed is initialised with world_size > 1."""
    steps = models[0].config.steps if steps is None else steps
    if _traced(models):
        return _run_traced(models, steps)
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
