"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` for rendezvous/gather.

Ensembles shard by replica index blocks ``[r*S/G, (r+1)*S/G)`` with no data-path collective
(SURVEY.md section 8(e)); the only communication is the final gather of ``[S/G][n_metrics]``
rows.  Every helper degrades to the identity when ``torch.distributed`` is not initialised.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def _td():
    try:
        import torch.distributed as td
        if td.is_available() and td.is_initialized():
            return td
    except Exception:
        pass
    return None


def rank_world() -> Tuple[int, int]:
    td = _td()
    return (td.get_rank(), td.get_world_size()) if td else (0, 1)


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of ``n`` items owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_range(n: int) -> Tuple[int, int]:
    r, w = rank_world()
    return shard_bounds(n, r, w)


def _device_for_backend(td):
    import torch
    if td.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def gather_rows(local: np.ndarray, n_total: int) -> np.ndarray:
    """All-gather of row blocks laid out by :func:`shard_bounds` -> the full ``[n_total, C]``."""
    td = _td()
    if td is None or td.get_world_size() == 1:
        return local
    import torch
    w = td.get_world_size()
    dev = _device_for_backend(td)
    cap = max(shard_bounds(n_total, r, w)[1] - shard_bounds(n_total, r, w)[0] for r in range(w))
    buf = torch.zeros((cap, local.shape[1]), dtype=torch.float64, device=dev)
    if local.shape[0]:
        buf[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    parts = [torch.zeros_like(buf) for _ in range(w)]
    td.all_gather(parts, buf)
    rows = []
    for r in range(w):
        lo, hi = shard_bounds(n_total, r, w)
        rows.append(parts[r][:hi - lo].cpu().numpy())
    return np.concatenate(rows, axis=0)


def max_over_ranks(value: float) -> float:
    td = _td()
    if td is None or td.get_world_size() == 1:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=_device_for_backend(td))
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float) -> float:
    td = _td()
    if td is None or td.get_world_size() == 1:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=_device_for_backend(td))
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return float(t.item())


def barrier() -> None:
    td = _td()
    if td is not None and td.get_world_size() > 1:
        td.barrier()
