"""Population sharding: ONE model split across the GPUs of a box, one process per GPU.

Well-mixed populations (agents coupled only through env scalars: C1, C4 -- SURVEY.md 8(e))
split by contiguous agent-index ranges; env is replicated.  The only exchange is the row of
per-step partial sums (<= 12 doubles) between the agent phase and the env phase.  It happens
INSIDE the step kernel over NVLink peer memory (``csrc/rules.cuh::peer_exchange``): every
rank exports the CUDA IPC handle of its exchange buffer, the handles are all-gathered through
``torch.distributed`` (plumbing only) and mapped once; after that no NCCL or host call sits on
the step path.  ``JXB_EXCHANGE=nccl`` selects the NCCL all-reduce arm instead (comparison).

Per-agent keys use the GLOBAL agent index (``split(key, N_global)[global_offset + i]``), so a
sharded run reproduces the unsharded model's state bit for bit; the reduction order differs
(per-rank partials folded in rank order), so float32 env trajectories agree to rounding.

Grids (C2, SURVEY.md 8(e) "Grid: row blocks + halo") split by ROW BANDS: rank r owns the rows
``shard_bounds(W, r, world)`` of the one ``env['grid']`` and a range of the empty-cell slots, and walks
only the movers of its own rows: a mover leaves as a 16-byte request stored into the receive area of the
rank that holds its slot, which rewrites the slot and forwards the mover to the owner of the target row
-- posted stores over NVLink and step flags only (``csrc/grid_shard.cuh``).  :func:`shard_model` on a Schelling model
selects it; state reads combine the ranks' views (``position`` max, ``satisfied`` min, ``moves``
sum) and are collective calls.  Results are bit-identical to the single-GPU run.

Networks (C3, SURVEY.md 8(e) "Network: 1-D node partition, replicated bit-packed state") split by NODE
RANGES at multiples of 32: rank r holds the CSR rows, the ``state`` slice and the draws of its agents
and a copy of the global "is infected" bitmap; the pull kernel stores the new bitmap word of every
32-row group straight into all ranks' copies (``csrc/sir.cuh``), so the neighbour aggregation is also
the all-gather of the state slices.  ``states['state']`` returns the rank's slice; the S/I/R counts
are exact integers folded in rank order, so the metric rows and states equal the single-GPU run's.  The ranks of a
sharded Grid / Network may share one device (``JXB_DEVICE=0`` for every rank, ``gloo`` process group): CUDA
IPC and the spin waits work across processes on the same GPU, which is how the parity tests cover
the whole path where only one GPU is visible.
"""
from __future__ import annotations

import os

import numpy as np

from . import _native as nat
from . import dist

_attached = False


def attach_peers() -> None:
    """Map every rank's exchange buffer into this process (idempotent)."""
    global _attached
    if _attached:
        return
    td = dist._td()
    if td is None or td.get_world_size() == 1:
        raise RuntimeError("population sharding needs torch.distributed initialised with world_size > 1")
    import torch
    rank, world = td.get_rank(), td.get_world_size()
    eng = nat.engine()
    lib = nat.lib()
    use_nccl = os.environ.get("JXB_EXCHANGE", "") == "nccl"
    dev = dist._device_for_backend(td)
    # NCCL communicator of the engine: the comparison arm of the scalar exchange (JXB_EXCHANGE=nccl)
    # and the bulk all-reduce of the economy's Gini histogram
    # (only with one GPU per rank: NCCL refuses two ranks on one device; ranks that time-share a GPU -- the
    # single-GPU parity tests, gloo process group -- run on the peer-memory exchange alone)
    if td.get_backend() == "nccl":
        ident = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            nat.check(lib.jxb_nccl_unique_id(nat.ptr(ident), ident.nbytes))
        t = torch.from_numpy(ident).to(dev)
        td.broadcast(t, src=0)
        ident = np.ascontiguousarray(t.cpu().numpy())
        nat.check(lib.jxb_engine_attach_nccl(eng.handle, nat.ptr(ident), ident.nbytes, rank, world))
    elif use_nccl:
        raise RuntimeError("JXB_EXCHANGE=nccl needs the nccl process group (one GPU per rank)")
    if not use_nccl:
        h = np.zeros(64, dtype=np.uint8)
        nat.check(lib.jxb_engine_p2p_export(eng.handle, nat.ptr(h), h.nbytes))
        parts = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
        td.all_gather(parts, torch.from_numpy(h).to(dev))
        table = np.ascontiguousarray(np.stack([p.cpu().numpy() for p in parts], axis=0))
        nat.check(lib.jxb_engine_p2p_attach(eng.handle, nat.ptr(table), 64, rank, world))
    td.barrier()
    _attached = True


class DistGroup:
    """The ranks of the ``torch.distributed`` process group (one process per GPU)."""
    def __init__(self):
        td = dist._td()
        if td is None or td.get_world_size() == 1:
            raise RuntimeError("population sharding needs torch.distributed initialised with world_size > 1")
        self._td = td
        self.rank, self.world = td.get_rank(), td.get_world_size()

    def all_gather_bytes(self, b: np.ndarray) -> np.ndarray:
        """uint8[n] of every rank -> uint8[world, n]."""
        import torch
        dev = dist._device_for_backend(self._td)
        mine = torch.from_numpy(np.ascontiguousarray(b, dtype=np.uint8)).to(dev)
        parts = [torch.zeros_like(mine) for _ in range(self.world)]
        self._td.all_gather(parts, mine)
        return np.stack([p.cpu().numpy() for p in parts], axis=0)

    def all_reduce(self, a: np.ndarray, op: str) -> np.ndarray:
        import torch
        td = self._td
        dev = dist._device_for_backend(td)
        src = np.ascontiguousarray(a)
        view = src.view(np.uint8) if src.dtype == np.bool_ else src
        t = torch.from_numpy(view).to(dev)
        td.all_reduce(t, op={"max": td.ReduceOp.MAX, "min": td.ReduceOp.MIN, "sum": td.ReduceOp.SUM}[op])
        out = t.cpu().numpy()
        return out.view(np.bool_) if src.dtype == np.bool_ else out

    def barrier(self) -> None:
        self._td.barrier()


def shard_model(model) -> None:
    """Mark an un-initialised core ``Model`` as sharded over the ranks of the process group:
    ``initialize()`` then allocates only this rank's index range of every collection (well-mixed
    populations), this rank's row band of the Grid (Schelling) or its node range of the Network (SIR)."""
    if model._is_initialized:
        raise RuntimeError("shard_model must be called before Model.initialize()")
    from .model import program_of
    # a Grid (row bands) and a Network (node ranges) bring their own peer-mapped receive areas
    own_areas = (program_of(model._update_state_fn, model._metrics_fn) in ("schelling", "sir")
                 and not model._needs_tracing())
    if not own_areas:
        attach_peers()            # engine-level exchange buffer + NCCL communicator of the well-mixed programs
    group = DistGroup()
    model._shard = (group.rank, group.world)
    model._shard_group = group


def network_cuts(n: int, edges, world: int, row_cost: int = 4, balance: str = None):
    """Node-range boundaries ``[0 = c_0 < c_1 < ... < c_world = n]`` of a sharded Network, at multiples
    of 32 (whole words of the infected bitmap).  ``balance="nodes"`` (default, or ``JXB_NET_SPLIT``)
    gives every rank the same number of agents; ``"entries"`` balances ``sum(out_degree + row_cost)``
    instead -- on a scale-free graph whose hubs sit at low indices an even split leaves the first rank
    with ~70 % of the adjacency at 2 ranks.  Measured at C3 on 2 B200 the even split is still the faster
    one (143 vs 169 us/step): the pull kernel's step time is set by its longest lane-serial row walks, not
    by the adjacency volume.  Every rank computes the same cuts from the same edge list."""
    groups = (n + 31) // 32
    balance = balance or os.environ.get("JXB_NET_SPLIT", "nodes")
    if edges is None or balance == "nodes" or groups < world:
        return [min(dist.shard_bounds(groups, r, world)[0] * 32, n) for r in range(world)] + [n]
    src = np.asarray(edges, dtype=np.int64).reshape(-1, 2)[:, 0]
    cost = np.bincount(src, minlength=n)[:n].astype(np.int64) + row_cost
    cum = np.cumsum(np.add.reduceat(cost, np.arange(0, n, 32)))
    cuts = [0]
    for r in range(1, world):
        g = int(np.searchsorted(cum, cum[-1] * r / world)) + 1
        g = max(g, cuts[-1] // 32 + 1)                 # at least one group per rank ...
        g = min(g, groups - (world - r))               # ... and room for the ranks that follow
        cuts.append(g * 32)
    return cuts + [n]


def local_edges(edges, lo: int, hi: int) -> np.ndarray:
    """The rows of ``env['network_edges']`` (``jaxabm/agentpy.py:557``) a node range owns: edges whose SOURCE lies
    in ``[lo, hi)``, sources re-based to local rows, targets left as global ids (what
    ``jxb_model_set_network`` takes on a sharded Network)."""
    e = np.asarray(edges, dtype=np.int32).reshape(-1, 2)
    e = e[(e[:, 0] >= lo) & (e[:, 0] < hi)].copy()
    e[:, 0] -= lo
    return e


def local_range(n: int):
    r, w = dist.rank_world()
    return dist.shard_bounds(n, r, w)
