"""ctypes binding of ``libjxb.so`` (C ABI declared in ``include/jxb.h``).

The library is the product: if it is missing, or the box has no CUDA device, every
call that needs it raises -- there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libjxb.so")

MAX_TYPES = 4
MAX_PARAMS = 16
MAX_METRICS = 32

RNG_LEGACY, RNG_PARTITIONABLE = 0, 1

RULE = dict(random_walker=1, scaled_walker=2, consumer=3, producer=4, growth=5, increment=6,
            wealth=7, schelling=8, sir=9, household=10, consumer_firm=11, traced=12)
PROGRAM = dict(none=0, random_walk=1, market=2, growth=3, counter=4, schelling=5, sir=6, economy=7, traced=8)

DTYPES = {0: np.float32, 1: np.int32, 2: np.bool_, 3: np.float64}


class JxbError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libjxb error {code}: {message}")
        self.code = code
        self.message = message


class TypeDesc(C.Structure):
    _fields_ = [("rule", C.c_int32), ("n_agents", C.c_int64), ("global_offset", C.c_int64),
                ("global_n", C.c_int64), ("n_params", C.c_int32), ("params", C.c_float * MAX_PARAMS)]


class ModelDesc(C.Structure):
    _fields_ = [("program", C.c_int32), ("rng_mode", C.c_int32), ("n_types", C.c_int32),
                ("types", TypeDesc * MAX_TYPES), ("n_params", C.c_int32),
                ("params", C.c_double * MAX_PARAMS), ("grid_w", C.c_int32), ("grid_h", C.c_int32),
                ("grid_periodic", C.c_int32), ("world_size", C.c_int32), ("rank", C.c_int32)]


class PredIns(C.Structure):
    """One instruction of a filter predicate program (``jxb_pred_ins``)."""
    _fields_ = [("op", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("f", C.c_float)]


_lib: Optional[C.CDLL] = None

# name -> (restype, argtypes); every symbol declared in include/jxb.h
_P = C.c_void_p
IPC_HANDLE_BYTES = 64        # include/jxb.h: JXB_IPC_HANDLE_BYTES
GRID_HANDLE_BYTES = 80       # include/jxb.h: JXB_GRID_HANDLE_BYTES (IPC handle + band + launch shape)

SIGNATURES = {
    "jxb_version": (C.c_int, []),
    "jxb_last_error": (C.c_char_p, []),
    "jxb_engine_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "jxb_engine_destroy": (C.c_int, [_P]),
    "jxb_engine_sm_count": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "jxb_engine_launch_count": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "jxb_model_create": (C.c_int, [_P, C.POINTER(ModelDesc), C.POINTER(_P)]),
    "jxb_model_create_traced": (C.c_int, [_P, C.POINTER(ModelDesc), _P, C.POINTER(_P)]),
    "jxb_model_destroy": (C.c_int, [_P]),
    "jxb_model_n_fields": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int)]),
    "jxb_model_field_info": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int),
                                       C.POINTER(C.c_int)]),
    "jxb_model_n_env": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "jxb_model_env_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int)]),
    "jxb_model_n_metrics": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "jxb_model_metric_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int)]),
    "jxb_model_upload": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_size_t]),
    "jxb_model_download": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_size_t]),
    "jxb_model_fill": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_size_t]),
    "jxb_model_set_env": (C.c_int, [_P, C.c_int, C.c_double]),
    "jxb_model_get_env": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
    "jxb_model_set_type_param": (C.c_int, [_P, C.c_int, C.c_int, C.c_float]),
    "jxb_model_set_network": (C.c_int, [_P, _P, C.c_int64]),
    "jxb_model_grid_rebuild": (C.c_int, [_P]),
    "jxb_model_download_grid": (C.c_int, [_P, _P, C.c_size_t]),
    "jxb_model_download_empty_cells": (C.c_int, [_P, _P, C.c_size_t]),
    "jxb_model_init": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "jxb_collection_init": (C.c_int, [_P, C.c_int, C.c_uint32, C.c_uint32]),
    "jxb_collection_update": (C.c_int, [_P, C.c_int, C.c_uint32, C.c_uint32]),
    "jxb_model_run": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "jxb_model_time_step": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "jxb_model_set_profile": (C.c_int, [_P, C.c_int]),
    "jxb_model_profile": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_char_p)]),
    "jxb_ensemble_run": (C.c_int, [_P, C.POINTER(ModelDesc), C.c_int, C.c_int, _P, _P, _P, _P, C.c_int, _P,
                                   C.POINTER(C.c_double)]),
    "jxb_nccl_unique_id": (C.c_int, [_P, C.c_size_t]),
    "jxb_engine_attach_nccl": (C.c_int, [_P, _P, C.c_size_t, C.c_int, C.c_int]),
    "jxb_engine_p2p_export": (C.c_int, [_P, _P, C.c_size_t]),
    "jxb_engine_p2p_attach": (C.c_int, [_P, _P, C.c_size_t, C.c_int, C.c_int]),
    "jxb_model_grid_shard_export": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_size_t]),
    "jxb_model_grid_shard_attach": (C.c_int, [_P, _P, C.c_size_t, C.c_int]),
    "jxb_model_net_shard_export": (C.c_int, [_P, _P, C.c_size_t]),
    "jxb_model_net_shard_attach": (C.c_int, [_P, _P, C.c_size_t, C.c_int]),
    "jxb_model_net_shard_sync": (C.c_int, [_P]),
    "jxb_model_record_fields": (C.c_int, [_P, C.c_int, _P, _P]),
    "jxb_model_series_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "jxb_model_series_download": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "jxb_collection_filter_select": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, C.c_size_t, C.POINTER(C.c_int64)]),
    "jxb_collection_filter_gather": (C.c_int, [_P, C.c_int, _P, C.c_int]),
    "jxb_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "jxb_host_free": (C.c_int, [_P]),
    "jxb_prng_split": (C.c_int, [C.c_int, _P, C.c_int, _P]),
    "jxb_prng_bits": (C.c_int, [C.c_int, _P, C.c_int64, _P]),
    "jxb_prng_uniform": (C.c_int, [C.c_int, _P, C.c_int64, C.c_float, C.c_float, _P]),
    "jxb_prng_randint": (C.c_int, [C.c_int, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    "jxb_prng_feistel": (C.c_int, [C.c_uint32, _P, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]),
    "jxb_prng_threefry2x32": (C.c_int, [_P, _P, _P]),
}


def lib() -> C.CDLL:
    """Load ``libjxb.so`` (built in-tree by ``__graft_entry__.build()``); fail loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(jaxabm_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)      # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise JxbError(rc, lib().jxb_last_error().decode("utf-8", "replace"))


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class _PinnedBlock:
    """Owner of one page-locked host block; NumPy arrays made from it keep it alive as their base."""

    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        check(lib().jxb_host_alloc(max(int(nbytes), 1), C.byref(self.ptr)))
        self.nbytes = int(nbytes)
        self.__array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr.value, False),
                                    "version": 3}

    def __del__(self):
        try:
            if self.ptr:
                lib().jxb_host_free(self.ptr)
                self.ptr = C.c_void_p()
        except Exception:
            pass


PINNED_MIN_BYTES = 1 << 20


def result_empty(shape, dtype) -> np.ndarray:
    """Uninitialised host array for a device read-back: page-locked (cached by the library) when it
    is at least 1 MB, plain NumPy memory otherwise."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    if n < PINNED_MIN_BYTES:
        return np.empty(shape, dtype=dtype)
    return np.asarray(_PinnedBlock(n)).view(dtype).reshape(shape)


_engine = None


class Engine:
    """One per process, bound to one device (``LOCAL_RANK`` under torchrun, else 0)."""

    def __init__(self, device: Optional[int] = None):
        if device is None:
            device = int(os.environ.get("JXB_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        self.device = device
        self.handle = C.c_void_p()
        check(lib().jxb_engine_create(device, C.byref(self.handle)))

    @property
    def sm_count(self) -> int:
        v = C.c_int()
        check(lib().jxb_engine_sm_count(self.handle, C.byref(v)))
        return v.value

    @property
    def launch_count(self) -> int:
        v = C.c_int64()
        check(lib().jxb_engine_launch_count(self.handle, C.byref(v)))
        return v.value


def engine() -> Engine:
    global _engine
    if _engine is None:
        _engine = Engine()
    return _engine


def default_rng_mode() -> int:
    """Stream layout of ``jax.random``: partitionable (JAX >= 0.5.0 default, what a fresh
    install of the reference's ``jax>=0.4.1`` pin resolves to) unless ``JXB_RNG_MODE=legacy``."""
    v = os.environ.get("JXB_RNG_MODE", "partitionable").lower()
    return RNG_LEGACY if v in ("legacy", "0", "original") else RNG_PARTITIONABLE
