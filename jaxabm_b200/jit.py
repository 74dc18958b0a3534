"""Compile a traced model (``trace.py``) into an in-tree shared library and hand it to ``libjxb``.

``nvcc -gencode arch=compute_100a,code=sm_100a`` with the flags of the main library
(``-fmad=false``: a*b+c rounds twice, as the unfused XLA ops do); the result is cached under
``jaxabm_b200/csrc/_jit/<sha1 of the source>.so``.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import shutil
import subprocess
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _native as nat
from . import trace as T

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
JIT_DIR = os.path.join(CSRC, "_jit")
_loaded: Dict[str, C.CDLL] = {}


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: traced rules are compiled at model build time")


NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]
_abi_salt: Optional[str] = None


def _abi() -> str:
    """Everything besides the generated source that decides the binary and its ABI towards libjxb: the headers the
    source includes (ModelDev is passed by value), the public header, the flags and the library version."""
    global _abi_salt
    if _abi_salt is None:
        h = hashlib.sha1()
        for name in sorted(os.listdir(CSRC)):
            if name.endswith(".cuh"):
                with open(os.path.join(CSRC, name), "rb") as f:
                    h.update(name.encode() + b"\0" + f.read())
        with open(os.path.join(os.path.dirname(_HERE), "include", "jxb.h"), "rb") as f:
            h.update(f.read())
        h.update(" ".join(NVCC_FLAGS).encode())
        h.update(str(nat.lib().jxb_version()).encode())
        _abi_salt = h.hexdigest()
    return _abi_salt


def compile_source(src: str) -> C.CDLL:
    h = hashlib.sha1((_abi() + "\0" + src).encode()).hexdigest()[:16]
    if h in _loaded:
        return _loaded[h]
    os.makedirs(JIT_DIR, exist_ok=True)
    so = os.path.join(JIT_DIR, f"jxc_{h}.so")
    if not os.path.exists(so):
        # per-process temporaries + one atomic rename: the ranks of a sharded traced ensemble all compile the
        # same source at the same time
        tag = f"{h}.{os.getpid()}"
        cu = os.path.join(JIT_DIR, f"jxc_{tag}.tmp.cu")
        tmp = os.path.join(JIT_DIR, f"jxc_{tag}.tmp.so")
        with open(cu, "w") as f:
            f.write(src)
        cmd = [_nvcc(), *NVCC_FLAGS, "-I", CSRC, cu, "-o", tmp]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise T.TraceError("the generated kernel did not compile (this is a tracer bug):\n" + r.stderr[-3000:])
        os.replace(cu, os.path.join(JIT_DIR, f"jxc_{h}.cu"))
        os.replace(tmp, so)
    lib = C.CDLL(so)
    _loaded[h] = lib
    return lib


def trace_variants(model) -> List[T.TracedModel]:
    """Trace with the initial (Python-typed) env, then with the env dtypes update_state_fn produces,
    until a trace's outputs have the dtypes of its inputs (normally two variants: the first step, in
    which env entries are still Python scalars, and every later step)."""
    variants: List[T.TracedModel] = []
    dts: Optional[Dict[str, str]] = None
    for _ in range(4):
        tm = T.trace_model(model._agent_collections, model._env_state, model._params, model._update_state_fn,
                           model._metrics_fn, model.config, dts)
        in_sig = dict(zip(tm.env_names, tm.env_dtypes))
        out_sig = dict(in_sig)
        out_sig.update({k: v.dtype for k, v in tm.env_out.items()})
        variants.append(tm)
        if out_sig == in_sig:
            break
        dts = out_sig
    else:
        raise T.TraceError("the env dtypes do not settle after three steps")
    if len(variants) > 2:
        raise T.TraceError("the env dtypes change more than once over the first steps; give the env entries their "
                           "final dtype (e.g. np.float32(...)) in add_env_state")
    if len(variants) == 2 and variants[0].env_names != variants[1].env_names:
        # keys added by update_state_fn: both variants must use the slot order of the later one
        variants[0] = T.trace_model(model._agent_collections, model._env_state, model._params, model._update_state_fn,
                                    model._metrics_fn, model.config, None, slot_order=variants[1].env_names)
    return variants


class TracedSpec(C.Structure):
    _fields_ = [("n_fields", C.c_int32 * nat.MAX_TYPES),
                ("field_names", (C.c_char_p * 20) * nat.MAX_TYPES),
                ("field_dtypes", (C.c_int32 * 20) * nat.MAX_TYPES),
                ("n_env", C.c_int32), ("env_names", C.c_char_p * 32), ("env_dtypes", C.c_int32 * 32),
                ("env_init", C.c_double * 32),
                ("n_metrics", C.c_int32), ("metric_names", C.c_char_p * nat.MAX_METRICS),
                ("metric_dtypes", C.c_int32 * nat.MAX_METRICS),
                ("has_env_fn", C.c_int32), ("n_acc", C.c_int32), ("n_variants", C.c_int32),
                ("launch_init", C.c_void_p), ("launch_step", C.c_void_p),
                ("n_consts", C.c_int32), ("consts", C.POINTER(C.c_double)),
                ("field_widths", (C.c_int32 * 20) * nat.MAX_TYPES)]


def build(model) -> Tuple[TracedSpec, C.CDLL, List[T.TracedModel], str]:
    variants = trace_variants(model)
    src, meta = T.generate_source(variants)
    lib = compile_source(src)
    tm = variants[-1]
    spec = TracedSpec()
    if len(tm.types) > nat.MAX_TYPES:
        raise T.TraceError(f"at most {nat.MAX_TYPES} agent collections per model")
    for i, t in enumerate(tm.types):
        if len(t["fields"]) > 20:
            raise T.TraceError("at most 20 state fields per collection")
        spec.n_fields[i] = len(t["fields"])
        for f, (name, dt, width) in enumerate(t["fields"]):
            spec.field_names[i][f] = name.encode()
            spec.field_dtypes[i][f] = T._DT_CODE[dt]
            spec.field_widths[i][f] = width
    if len(tm.env_names) > 32 or len(tm.metrics) > nat.MAX_METRICS:
        raise T.TraceError("at most 32 scalar env entries and 32 metrics")
    spec.n_env = len(tm.env_names)
    v0 = variants[0]
    for k, name in enumerate(tm.env_names):
        spec.env_names[k] = name.encode()
        spec.env_dtypes[k] = T._DT_CODE[tm.env_dtypes[k]]
        init = model._env_state.get(name, 0.0)
        spec.env_init[k] = float(init) if T._env_dtype_of(init) is not None else 0.0
    spec.n_metrics = len(tm.metrics)
    for k, (name, v) in enumerate(tm.metrics):
        spec.metric_names[k] = name.encode()
        spec.metric_dtypes[k] = T._DT_CODE[v.dtype]
    spec.has_env_fn = 1 if tm.has_env_fn else 0
    spec.n_acc = int(lib.jxc_n_acc())
    spec.n_variants = int(lib.jxc_n_variants())
    spec.launch_init = C.cast(lib.jxc_launch_init, C.c_void_p)
    spec.launch_step = C.cast(lib.jxc_launch_step, C.c_void_p)
    consts = (C.c_double * max(len(meta["consts"]), 1))(*meta["consts"])
    spec.n_consts = len(meta["consts"])
    spec.consts = C.cast(consts, C.POINTER(C.c_double))
    spec._keep_consts = consts
    return spec, lib, variants, src
