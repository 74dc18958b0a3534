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
"""
from __future__ import annotations

import ctypes as C
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
    ident = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        nat.check(lib.jxb_nccl_unique_id(nat.ptr(ident), ident.nbytes))
    t = torch.from_numpy(ident).to(dev)
    td.broadcast(t, src=0)
    ident = np.ascontiguousarray(t.cpu().numpy())
    nat.check(lib.jxb_engine_attach_nccl(eng.handle, nat.ptr(ident), ident.nbytes, rank, world))
    if not use_nccl:
        h = np.zeros(64, dtype=np.uint8)
        nat.check(lib.jxb_engine_p2p_export(eng.handle, nat.ptr(h), h.nbytes))
        parts = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
        td.all_gather(parts, torch.from_numpy(h).to(dev))
        table = np.ascontiguousarray(np.stack([p.cpu().numpy() for p in parts], axis=0))
        nat.check(lib.jxb_engine_p2p_attach(eng.handle, nat.ptr(table), 64, rank, world))
    td.barrier()
    _attached = True


def shard_model(model) -> None:
    """Mark an un-initialised core ``Model`` as sharded over the ranks of the process group:
    ``initialize()`` then allocates only this rank's index range of every collection."""
    if model._is_initialized:
        raise RuntimeError("shard_model must be called before Model.initialize()")
    attach_peers()
    model._shard = dist.rank_world()


def local_range(n: int):
    r, w = dist.rank_world()
    return dist.shard_bounds(n, r, w)
