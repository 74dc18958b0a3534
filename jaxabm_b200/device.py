"""Thin object wrapper over a ``jxb_model`` handle (device-resident model state)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native as nat


class TypeSpec:
    """One collection of a device model: rule, local/global population, rule constants."""

    def __init__(self, rule: str, n_agents: int, params: Sequence[float] = (),
                 global_offset: int = 0, global_n: Optional[int] = None):
        self.rule = rule
        self.n_agents = int(n_agents)
        self.params = [float(p) for p in params]
        self.global_offset = int(global_offset)
        self.global_n = int(global_n) if global_n is not None else int(n_agents)


def make_desc(program: str, types: Sequence[TypeSpec], params: Sequence[float] = (),
              rng_mode: Optional[int] = None, grid: Optional[Tuple[int, int, bool]] = None,
              world_size: int = 1, rank: int = 0) -> nat.ModelDesc:
    d = nat.ModelDesc()
    d.program = nat.PROGRAM[program]
    d.rng_mode = nat.default_rng_mode() if rng_mode is None else int(rng_mode)
    d.n_types = len(types)
    if len(types) > nat.MAX_TYPES:
        raise ValueError(f"at most {nat.MAX_TYPES} agent collections per model")
    for i, t in enumerate(types):
        td = d.types[i]
        td.rule = nat.RULE[t.rule]
        td.n_agents = t.n_agents
        td.global_offset = t.global_offset
        td.global_n = t.global_n
        td.n_params = len(t.params)
        for k, v in enumerate(t.params):
            td.params[k] = v
    d.n_params = len(params)
    for k, v in enumerate(params):
        d.params[k] = float(v)
    if grid is not None:
        d.grid_w, d.grid_h, d.grid_periodic = int(grid[0]), int(grid[1]), int(bool(grid[2]))
    d.world_size, d.rank = world_size, rank
    return d


class DeviceModel:
    def __init__(self, desc: nat.ModelDesc, engine: Optional[nat.Engine] = None, traced_spec=None, keep=None):
        self.engine = engine or nat.engine()
        self.desc = desc
        self.handle = C.c_void_p()
        self._lib = nat.lib()
        self._keep = keep                 # the generated library of a traced model must outlive the handle
        self.group = None                 # host-side rank group of a row-band sharded Grid
        self.band = None                  # (row_begin, row_end) of this rank
        self.net_group = None             # host-side rank group of a node-range sharded Network
        if traced_spec is not None:
            nat.check(self._lib.jxb_model_create_traced(self.engine.handle, C.byref(desc), C.byref(traced_spec),
                                                        C.byref(self.handle)))
        else:
            nat.check(self._lib.jxb_model_create(self.engine.handle, C.byref(desc), C.byref(self.handle)))
        self.n_types = desc.n_types
        self.fields: List[List[Tuple[str, np.dtype, int]]] = []
        for t in range(self.n_types):
            nf = C.c_int()
            nat.check(self._lib.jxb_model_n_fields(self.handle, t, C.byref(nf)))
            fl = []
            for f in range(nf.value):
                name, dt, w = C.c_char_p(), C.c_int(), C.c_int()
                nat.check(self._lib.jxb_model_field_info(self.handle, t, f, C.byref(name), C.byref(dt), C.byref(w)))
                fl.append((name.value.decode(), nat.DTYPES[dt.value], w.value))
            self.fields.append(fl)
        n = C.c_int()
        nat.check(self._lib.jxb_model_n_env(self.handle, C.byref(n)))
        self.env_slots: List[Tuple[str, np.dtype]] = []
        for s in range(n.value):
            name, dt = C.c_char_p(), C.c_int()
            nat.check(self._lib.jxb_model_env_info(self.handle, s, C.byref(name), C.byref(dt)))
            self.env_slots.append((name.value.decode(), nat.DTYPES[dt.value]))
        nat.check(self._lib.jxb_model_n_metrics(self.handle, C.byref(n)))
        self.metric_slots: List[Tuple[str, np.dtype]] = []
        for s in range(n.value):
            name, dt = C.c_char_p(), C.c_int()
            nat.check(self._lib.jxb_model_metric_info(self.handle, s, C.byref(name), C.byref(dt)))
            self.metric_slots.append((name.value.decode(), nat.DTYPES[dt.value]))

    def close(self):
        if self.handle:
            self._lib.jxb_model_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- fields -------------------------------------------------------------------
    def field_index(self, t: int, name: str) -> int:
        for i, (n, _, _) in enumerate(self.fields[t]):
            if n == name:
                return i
        raise KeyError(name)

    def n_agents(self, t: int) -> int:
        return int(self.desc.types[t].n_agents)

    def _shape(self, t: int, f: int):
        _, dt, w = self.fields[t][f]
        n = self.n_agents(t)
        return ((n,) if w == 1 else (n, w)), dt

    # how the ranks' views of a sharded Grid's columns combine (include/jxb.h, grid sharding)
    GRID_COMBINE = {"position": "max", "satisfied": "min", "moves": "sum"}

    def download(self, t: int, f: int, out: Optional[np.ndarray] = None, combine: bool = True) -> np.ndarray:
        """Copy one state column to the host (into ``out`` -- e.g. a pinned buffer -- if given).  On a
        row-band sharded Grid this is a collective call that returns the whole population's column
        (``combine=False``: this rank's view only)."""
        shape, dt = self._shape(t, f)
        if out is None:
            out = nat.result_empty(shape, dt)
        elif out.shape != shape or out.dtype != dt or not out.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous {dt} array of shape {shape}")
        nat.check(self._lib.jxb_model_download(self.handle, t, f, nat.ptr(out), out.nbytes))
        op = self.GRID_COMBINE.get(self.fields[t][f][0]) if (self.group is not None and combine) else None
        if op is not None:
            out[...] = self.group.all_reduce(out, op)
        return out

    def grid_shard_setup(self, group) -> None:
        """Row band of this rank + receive areas of all ranks (``jxb_model_grid_shard_export/attach``)."""
        from .dist import shard_bounds
        lo, hi = shard_bounds(int(self.desc.grid_w), group.rank, group.world)
        handle = np.zeros(nat.GRID_HANDLE_BYTES, dtype=np.uint8)          # IPC handle + the band
        nat.check(self._lib.jxb_model_grid_shard_export(self.handle, lo, hi, nat.ptr(handle), handle.nbytes))
        table = np.ascontiguousarray(group.all_gather_bytes(handle))
        nat.check(self._lib.jxb_model_grid_shard_attach(self.handle, nat.ptr(table), table.shape[1], group.world))
        group.barrier()
        self.group, self.band = group, (lo, hi)

    def net_shard_setup(self, group) -> None:
        """Receive areas (global infected bitmaps) of all ranks (``jxb_model_net_shard_export/attach``)."""
        handle = np.zeros(64, dtype=np.uint8)
        nat.check(self._lib.jxb_model_net_shard_export(self.handle, nat.ptr(handle), handle.nbytes))
        table = np.ascontiguousarray(group.all_gather_bytes(handle))
        nat.check(self._lib.jxb_model_net_shard_attach(self.handle, nat.ptr(table), table.shape[1], group.world))
        group.barrier()
        self.net_group = group

    def net_shard_sync(self) -> None:
        """Collective: every rank hands its slice of the packed infected bitmap to all peers.  Barriers on both
        sides: nobody writes into a peer's area while that peer may still be rebuilding it (set_network clears
        the area), and nobody steps before every copy is whole."""
        self.net_group.barrier()
        nat.check(self._lib.jxb_model_net_shard_sync(self.handle))
        self.net_group.barrier()

    def upload(self, t: int, f: int, value) -> None:
        shape, dt = self._shape(t, f)
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=dt), shape))
        name = self.fields[t][f][0]
        if self.group is not None and name in ("type", "position"):
            # a row-band sharded Grid (collective call): the write makes every rank rebuild its cell binning from its
            # API columns, which only hold the band's view once they were read -- hand every rank the whole
            # 'position' / 'moves' columns first (the agents' move counts travel with them into the rebuilt bands)
            for other in ("position", "moves"):
                if other != name:
                    fo = self.field_index(t, other)
                    whole = self.download(t, fo)
                    nat.check(self._lib.jxb_model_upload(self.handle, t, fo, nat.ptr(whole), whole.nbytes))
        nat.check(self._lib.jxb_model_upload(self.handle, t, f, nat.ptr(a), a.nbytes))
        if self.net_group is not None:
            self.net_shard_sync()

    def fill(self, t: int, f: int, value) -> None:
        _, dt, w = self.fields[t][f]
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=dt), (w,)))
        nat.check(self._lib.jxb_model_fill(self.handle, t, f, nat.ptr(a), a.nbytes))

    # ---- env ----------------------------------------------------------------------
    def env_index(self, name: str) -> Optional[int]:
        for i, (n, _) in enumerate(self.env_slots):
            if n == name:
                return i
        return None

    def set_env(self, slot: int, value) -> None:
        nat.check(self._lib.jxb_model_set_env(self.handle, slot, float(value)))

    def get_env(self, slot: int):
        v = C.c_double()
        nat.check(self._lib.jxb_model_get_env(self.handle, slot, C.byref(v)))
        dt = self.env_slots[slot][1]
        return float(v.value) if dt == np.float64 else dt(v.value)

    def set_type_param(self, t: int, index: int, value: float) -> None:
        nat.check(self._lib.jxb_model_set_type_param(self.handle, t, index, float(value)))

    # ---- structure ----------------------------------------------------------------
    def set_network(self, edges) -> None:
        e = np.ascontiguousarray(np.asarray(edges, dtype=np.int32).reshape(-1, 2))
        nat.check(self._lib.jxb_model_set_network(self.handle, nat.ptr(e), e.shape[0]))
        if self.net_group is not None:         # the areas were rebuilt: redistribute the bitmap slices
            self.net_shard_sync()

    def grid_rebuild(self) -> None:
        nat.check(self._lib.jxb_model_grid_rebuild(self.handle))

    def download_grid(self) -> np.ndarray:
        out = np.empty((self.desc.grid_w, self.desc.grid_h), dtype=np.int32)
        nat.check(self._lib.jxb_model_download_grid(self.handle, nat.ptr(out), out.nbytes))
        if self.group is not None:         # every rank contributes its own rows (cells are >= -1)
            lo, hi = self.band
            out[:lo] = -2
            out[hi:] = -2
            out = self.group.all_reduce(out, "max")
        return out

    def download_empty_cells(self) -> np.ndarray:
        """env['empty_cells'] as int32[e, 2] (x, y) pairs in slot order."""
        e = int(self.desc.grid_w) * int(self.desc.grid_h) - self.n_agents(0)
        ids = np.empty(e, dtype=np.int32)
        if self.group is not None:         # the slots live in the ranks' receive areas: every rank must be done stepping
            self.group.barrier()
        nat.check(self._lib.jxb_model_download_empty_cells(self.handle, nat.ptr(ids), ids.nbytes))
        h = int(self.desc.grid_h)
        return np.stack([ids // h, ids % h], axis=1).astype(np.int32)

    # ---- time loop ----------------------------------------------------------------
    def init(self, key) -> None:
        k = np.asarray(key, dtype=np.uint32).reshape(2)
        nat.check(self._lib.jxb_model_init(self.handle, int(k[0]), int(k[1])))

    def collection_init(self, t: int, key) -> None:
        k = np.asarray(key, dtype=np.uint32).reshape(2)
        nat.check(self._lib.jxb_collection_init(self.handle, t, int(k[0]), int(k[1])))

    def collection_update(self, t: int, key) -> None:
        k = np.asarray(key, dtype=np.uint32).reshape(2)
        nat.check(self._lib.jxb_collection_update(self.handle, t, int(k[0]), int(k[1])))

    @property
    def time_step(self) -> int:
        v = C.c_int64()
        nat.check(self._lib.jxb_model_time_step(self.handle, C.byref(v)))
        return v.value

    def run(self, steps: int, collect_interval: int = 1):
        """-> (record_steps int32[n], metrics float64[n, n_metrics], device_seconds)."""
        t0 = self.time_step
        n_rec = (t0 + steps) // collect_interval - t0 // collect_interval
        m = np.zeros((max(n_rec, 1), nat.MAX_METRICS), dtype=np.float64)
        st = np.zeros(max(n_rec, 1), dtype=np.int32)
        n = C.c_int()
        secs = C.c_double()
        nat.check(self._lib.jxb_model_run(self.handle, int(steps), int(collect_interval), nat.ptr(m), nat.ptr(st),
                                          C.byref(n), C.byref(secs)))
        return st[:n.value], m[:n.value, :len(self.metric_slots)], secs.value

    # ---- record / select (include/jxb.h "record / select") -----------------------------
    def record_fields(self, pairs: Sequence[Tuple[int, int]]) -> None:
        """Snapshot the (collection, field) columns whenever a later run() records a history row."""
        types = np.ascontiguousarray([p[0] for p in pairs], dtype=np.int32)
        flds = np.ascontiguousarray([p[1] for p in pairs], dtype=np.int32)
        nat.check(self._lib.jxb_model_record_fields(self.handle, len(pairs), nat.ptr(types) if len(pairs) else None,
                                                    nat.ptr(flds) if len(pairs) else None))
        self._recorded = list(pairs)

    def series(self, k: int) -> np.ndarray:
        """Snapshots of recorded column k taken by the last run(): ``[n_records, N(, w)]``."""
        n, b = C.c_int(), C.c_size_t()
        nat.check(self._lib.jxb_model_series_info(self.handle, k, C.byref(n), C.byref(b)))
        t, f = self._recorded[k]
        shape, dt = self._shape(t, f)
        out = nat.result_empty((n.value,) + tuple(shape), dt)
        nat.check(self._lib.jxb_model_series_download(self.handle, k, nat.ptr(out), out.nbytes))
        return out

    def filter_select(self, t: int, program=None, mask: Optional[np.ndarray] = None) -> int:
        """Flag the agents of collection t (predicate program or host mask) -> number selected."""
        count = C.c_int64()
        if program is not None:
            arr = (nat.PredIns * len(program))(*[nat.PredIns(*ins) for ins in program])
            nat.check(self._lib.jxb_collection_filter_select(self.handle, t, arr, len(program), None, 0, C.byref(count)))
        else:
            m8 = np.ascontiguousarray(np.asarray(mask).astype(np.bool_).reshape(-1)).view(np.uint8)
            nat.check(self._lib.jxb_collection_filter_select(self.handle, t, None, 0, nat.ptr(m8), m8.nbytes, C.byref(count)))
        return int(count.value)

    def filter_gather(self, t: int, dst: "DeviceModel", dst_t: int = 0) -> None:
        nat.check(self._lib.jxb_collection_filter_gather(self.handle, t, dst.handle, dst_t))

    def set_profile(self, on: bool) -> None:
        nat.check(self._lib.jxb_model_set_profile(self.handle, int(on)))

    def profile(self):
        s, n, name = C.c_double(), C.c_int64(), C.c_char_p()
        nat.check(self._lib.jxb_model_profile(self.handle, C.byref(s), C.byref(n), C.byref(name)))
        return s.value, n.value, name.value.decode()


PROGRAM_ENV = {   # env slot names per program (csrc/engine.cu kPrograms)
    "none": [], "random_walk": ["bounds_lo", "bounds_hi", "time", "mean_x", "mean_y", "num_red", "num_blue"],
    "market": ["price_level", "gdp", "unemployment", "total_consumption", "total_production"],
    "growth": ["price_level", "interest_rate"], "counter": ["counter", "increment"],
    "schelling": ["segregation_index", "percent_satisfied", "total_moves"], "sir": [],
}
PROGRAM_N_METRICS = {0: 0, 1: 7, 2: 5, 3: 3, 4: 2, 5: 3, 6: 3, 7: 29}


def ensemble_run(desc: nat.ModelDesc, slots: Sequence[int], params: np.ndarray, seeds: np.ndarray,
                 steps: int, env_init: Optional[np.ndarray] = None, engine: Optional[nat.Engine] = None):
    """-> (last_metrics float64[R, n_metrics], device_seconds)."""
    eng = engine or nat.engine()
    seeds = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint32))
    R = seeds.shape[0]
    slots_a = np.ascontiguousarray(np.asarray(slots, dtype=np.int32))
    params = np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(R, len(slots_a)))
    n_m = PROGRAM_N_METRICS[int(desc.program)]
    out = np.zeros((R, max(n_m, 1)), dtype=np.float64)
    env = None if env_init is None else np.ascontiguousarray(np.asarray(env_init, dtype=np.float64))
    secs = C.c_double()
    nat.check(nat.lib().jxb_ensemble_run(eng.handle, C.byref(desc), R, len(slots_a), nat.ptr(slots_a),
                                         nat.ptr(params), nat.ptr(seeds),
                                         None if env is None else nat.ptr(env), int(steps), nat.ptr(out),
                                         C.byref(secs)))
    return out[:, :n_m], secs.value
