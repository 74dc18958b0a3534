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
        L.orc_sir_step.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_float, C.c_void_p]
        L.orc_walk_step.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                    C.POINTER(C.c_double), C.POINTER(C.c_float)]
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def use_all_cores() -> int:
    """OpenMP threads = the cores this process may run on, whatever OMP_NUM_THREADS says (torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would make the CPU baseline a single-thread number)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_num_threads(int(n))
    return num_threads()


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


class MarketFast:
    """Consumer/producer market (C4-A) on the C oracle: same initial state as
    ``oracle.rules.create_economy_model(...)`` (incomes / capital from the per-agent keys of
    ``Model.initialize``), stepped by ``orc_market_step``."""

    def __init__(self, num_consumers, num_producers, base_income=1.0, propensity_to_consume=0.8,
                 initial_capital=10.0, productivity=1.0, reinvestment_rate=0.3, price_adjustment_rate=0.1,
                 seed=0, mode=1):
        self.nc, self.np_ = int(num_consumers), int(num_producers)
        self.ptc, self.prd, self.rr, self.rate = (np.float32(propensity_to_consume), np.float32(productivity),
                                                  np.float32(reinvestment_rate), np.float32(price_adjustment_rate))
        keys = jl.split(jl.PRNGKey(seed), 3, mode)                  # model.py:129-137
        u_c = jl.uniform_scalar_batched(jl.split(keys[1], self.nc, mode), mode=mode)
        u_p = jl.uniform_scalar_batched(jl.split(keys[2], self.np_, mode), mode=mode)
        f32 = np.float32
        self.income = (f32(base_income) * (f32(0.8) + f32(0.4) * u_c)).astype(f32)
        self.capital = (f32(initial_capital) * (f32(0.8) + f32(0.4) * u_p)).astype(f32)
        self.savings = np.zeros(self.nc, f32)
        self.consumption = np.zeros(self.nc, f32)
        self.utility = np.zeros(self.nc, f32)
        self.production = np.zeros(self.np_, f32)
        self.profit = np.zeros(self.np_, f32)
        self.env = np.array([1.0, 0.0, 0.0, 0.0, 0.0], dtype=f32)

    def run(self, steps: int):
        out = {"gdp": [], "price_level": [], "unemployment": [], "avg_utility": [], "avg_profit": []}
        su, sp = C.c_double(), C.c_double()
        for _ in range(steps):
            lib().orc_market_step(self.nc, _p(self.savings), _p(self.consumption), _p(self.utility), _p(self.income),
                                  C.c_float(self.ptc), self.np_, _p(self.capital), _p(self.production),
                                  _p(self.profit), C.c_float(self.prd), C.c_float(self.rr), C.c_float(self.rate),
                                  _p(self.env), C.byref(su), C.byref(sp))
            out["price_level"].append(np.float32(self.env[0]))
            out["gdp"].append(np.float32(self.env[1]))
            out["unemployment"].append(np.float32(self.env[2]))
            out["avg_utility"].append(np.float32(su.value / max(self.nc, 1)))
            out["avg_profit"].append(np.float32(sp.value / max(self.np_, 1)))
        return out


class SirFast:
    """SIR on a network (C3) on the C oracle: same inputs and outputs as
    ``oracle.rules.create_sir_model(...).run(steps)`` (no env function: the key schedule of
    ``model.py:156,164`` only)."""

    def __init__(self, n, edges, beta=0.05, gamma=0.1, initial_infected=0.01, seed=42, mode=1):
        from .rules import edges_to_csr, sir_escape_table
        self.n, self.mode, self.gamma = int(n), mode, np.float32(gamma)
        rp, col = edges_to_csr(self.n, edges)
        self.row_ptr = np.ascontiguousarray(rp, dtype=np.int64)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.q = np.ascontiguousarray(sir_escape_table(beta), dtype=np.float32)
        init_keys, _, _, rng = key_schedule(seed, 1, False, 0, mode)
        u = jl.uniform_scalar_batched(jl.split(init_keys[0], self.n, mode), mode=mode)
        self.state = np.ascontiguousarray((u < np.float32(initial_infected)).astype(np.int32))
        self._next = np.empty_like(self.state)
        self._rng = rng

    def run(self, steps: int):
        coll = np.zeros((steps, 1, 2), dtype=np.uint32)
        upd = np.zeros((steps, 2), dtype=np.uint32)
        lib().orc_key_schedule(self.mode, _p(self._rng), 1, 0, steps, _p(coll), _p(upd))
        out = {"count_S": [], "count_I": [], "count_R": []}
        counts = np.zeros(3, dtype=np.int64)
        for t in range(steps):
            ck = np.ascontiguousarray(coll[t, 0])
            lib().orc_sir_step(self.mode, _p(ck), self.n, _p(self.row_ptr), _p(self.col), _p(self.state), _p(self._next),
                               _p(self.q), C.c_float(self.gamma), _p(counts))
            self.state, self._next = self._next, self.state
            for k, c in zip(("count_S", "count_I", "count_R"), counts):
                out[k].append(np.int32(c))
        return out


class WalkFast:
    """Random walkers (C1, ``examples/basic_example.py:33-69,141-182``) on the C oracle, from given
    ``position`` / ``velocity`` columns (the scaled bench variant draws them per agent)."""

    def __init__(self, position, velocity, bounds=(0.0, 1.0)):
        self.pos = np.ascontiguousarray(position, dtype=np.float32).copy()
        self.vel = np.ascontiguousarray(velocity, dtype=np.float32).copy()
        self.n = self.pos.shape[0]
        self.color = np.zeros(self.n, dtype=np.int32)
        self.steps_taken = np.zeros(self.n, dtype=np.int32)
        self.lo, self.hi = np.float32(bounds[0]), np.float32(bounds[1])

    def run(self, steps: int):
        out = {"mean_distance": [], "max_distance": []}
        sd, md = C.c_double(), C.c_float()
        for _ in range(steps):
            lib().orc_walk_step(self.n, _p(self.pos), _p(self.vel), _p(self.color), _p(self.steps_taken),
                                C.c_float(self.lo), C.c_float(self.hi), C.byref(sd), C.byref(md))
            out["mean_distance"].append(np.float32(sd.value / self.n))
            out["max_distance"].append(np.float32(md.value))
        return out
