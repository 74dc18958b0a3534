"""Seeded synthetic workload inputs (host side; used by bench.py and the tests)."""
from __future__ import annotations

import numpy as np


def scale_free_edges(n: int, m: int, seed: int) -> np.ndarray:
    """Barabasi-Albert-style undirected graph, vectorised (edge-copy construction).

    Node v >= 1 adds ``m`` edges; edge j (source ``j // m + 1``) picks its target as a
    uniformly random endpoint among the half-edges created before it -- which is
    preferential attachment.  References to a *target* slot of an earlier edge are chased
    by pointer jumping (they always point backwards, so every pass resolves the earliest
    open reference at least).  Returns the directed adjacency list int32[2E', 2] (both
    directions, as ``Network(directed=False)`` stores it); self-loops are dropped.
    """
    rng = np.random.RandomState(seed)
    E = (n - 1) * m
    j = np.arange(E, dtype=np.int64)
    src = j // m + 1
    r = (rng.random_sample(E) * (2 * j + 1)).astype(np.int64)      # slot in [0, 2j]
    r = np.minimum(r, 2 * j)
    tgt = np.full(E, -1, dtype=np.int64)
    own = r == 2 * j                       # the reserved slot: attach to node 0's seed stub
    tgt[own] = 0
    even = (~own) & (r % 2 == 0)           # slot 2i   -> source of edge i
    tgt[even] = r[even] // 2 // m + 1
    ref = np.where((~own) & (r % 2 == 1), r // 2, -1)             # slot 2i+1 -> target of edge i
    pend = np.nonzero(ref >= 0)[0]
    while pend.size:
        t = tgt[ref[pend]]
        done = t >= 0
        tgt[pend[done]] = t[done]
        pend = pend[~done]
    keep = src != tgt
    a, b = src[keep].astype(np.int32), tgt[keep].astype(np.int32)
    return np.concatenate([np.stack([a, b], 1), np.stack([b, a], 1)], axis=0)


def ring_lattice_edges(n: int, k: int) -> np.ndarray:
    """Each node linked to its k nearest neighbours on each side (both directions)."""
    i = np.arange(n, dtype=np.int64)
    parts = []
    for d in range(1, k + 1):
        parts.append(np.stack([i, (i + d) % n], 1))
        parts.append(np.stack([(i + d) % n, i], 1))
    return np.concatenate(parts, axis=0).astype(np.int32)
