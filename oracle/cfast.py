"""ctypes front end of ``oracle/c/liboracle.so`` (C/OpenMP restatement; TEST INFRASTRUCTURE
ONLY -- see ``oracle/__init__.py``).  Used for the multi-core CPU baseline in ``bench.py`` and
for parity checks at sizes the NumPy oracle cannot reach.  Validated against the NumPy
oracle in ``tests/test_oracle.py``."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import jaxlike as jl

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_LIB = os.path.join(_DIR, "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            subprocess.check_call(["make", "-s", "-C", _DIR])
        L = C.CDLL(_LIB)
        L.orc_num_threads.restype = C.c_int
        L.orc_schelling_step.restype = C.c_int64
        L.orc_schelling_step.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_double),
                                         C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_key_schedule.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_market_step.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                      C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                      C.c_float, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def key_schedule(seed: int, n_collections: int, has_env_fn: bool, steps: int, mode: int):
    """Same contract as ``oracle.runtime.key_schedule`` but O(steps) in C."""
    rng = jl.PRNGKey(seed)
    keys = jl.split(rng, n_collections + 1, mode)
    rng = np.ascontiguousarray(keys[0])
    coll = np.zeros((steps, n_collections, 2), dtype=np.uint32)
    upd = np.zeros((steps, 2), dtype=np.uint32)
    lib().orc_key_schedule(mode, _p(rng), n_collections, int(has_env_fn), steps, _p(coll), _p(upd))
    return keys[1:].copy(), coll, upd, rng


class SchellingFast:
    """Whole Schelling model on the C oracle: same inputs and outputs as
    ``oracle.rules.create_schelling_model(...).run(steps)``."""

    def __init__(self, grid_size, types, positions, similarity_threshold=0.5, periodic=False, seed=42, mode=1):
        self.G, self.mode, self.seed = int(grid_size), mode, seed
        self.thr, self.periodic = float(similarity_threshold), int(bool(periodic))
        self.type = np.ascontiguousarray(types, dtype=np.int32)
        self.pos = np.ascontiguousarray(positions, dtype=np.int32).copy()
        self.n = self.type.shape[0]
        self.satisfied = np.zeros(self.n, dtype=np.uint8)
        self.moves = np.zeros(self.n, dtype=np.int32)
        self.grid = -np.ones(self.G * self.G, dtype=np.int32)
        self.grid[self.pos[:, 0].astype(np.int64) * self.G + self.pos[:, 1]] = self.type
        self._cell_agent = np.empty(self.G * self.G, dtype=np.int32)
        self._U = np.empty(self.n, dtype=np.int32)
        self.E = np.ascontiguousarray(np.nonzero(self.grid < 0)[0].astype(np.int32))   # env['empty_cells']
        _, _, _, rng = key_schedule(seed, 1, True, 0, mode)
        self._rng = rng
        self.total_moves = 0
        self.time_step = 0

    def run(self, steps: int):
        coll = np.zeros((steps, 1, 2), dtype=np.uint32)
        upd = np.zeros((steps, 2), dtype=np.uint32)
        lib().orc_key_schedule(self.mode, _p(self._rng), 1, 1, steps, _p(coll), _p(upd))
        out = {"step": [], "percent_satisfied": [], "segregation_index": [], "total_moves": []}
        ssum, scnt, nu = C.c_double(), C.c_int64(), C.c_int64()
        for t in range(steps):
            ck = np.ascontiguousarray(coll[t, 0])
            m = lib().orc_schelling_step(self.G, self.G, self.periodic, self.thr, self.mode, _p(ck), self.n,
                                         _p(self.type), _p(self.pos), _p(self.satisfied), _p(self.moves),
                                         _p(self.grid), _p(self._cell_agent), _p(self._U), _p(self.E),
                                         self.E.shape[0], C.byref(ssum), C.byref(scnt), C.byref(nu))
            self.total_moves += int(m)
            self.time_step += 1
            out["step"].append(self.time_step)
            out["percent_satisfied"].append(np.float32((self.n - nu.value) / self.n))
            out["segregation_index"].append(np.float32(ssum.value / max(1, scnt.value)))
            out["total_moves"].append(np.int32(self.total_moves))
        return out
