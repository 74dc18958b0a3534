#!/usr/bin/env python
"""CPU model (not a measurement) of how the SIR pull kernel's work is spread over its warps.

`sir_pull_s_kernel` (csrc/sir.cuh) gives 32-row group g to warp g % W (W = 148 SMs x 8 CTAs x 8 warps) and a warp
walks its susceptible rows with dependent loads: a row of <= 256 entries by its own lane, two entries per
iteration; a longer row by all 32 lanes, 128 entries per iteration.  The script replays the seeded epidemic on
the C oracle and counts, per step, those dependent iterations per warp -- for the whole network and for the two
halves of an even 2-rank node split.  A kernel bound by latency finishes with its slowest warp, so `max warp`
(not `mean warp`) is what its time follows.

  python scripts/sir_warp_imbalance.py [N]      (N = 2,000,000 by default; 10,000,000 = C3)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from jaxabm_b200 import synthetic
from oracle import cfast

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
edges = synthetic.scale_free_edges(n, 5, 42)
f = cfast.SirFast(n, edges, beta=0.05, gamma=0.1, initial_infected=0.01, seed=42, mode=1)
deg = np.diff(f.row_ptr).astype(np.int64)
W = 148 * 8 * 8


def stats(state, lo, hi):
    d, s = deg[lo:hi], state[lo:hi]
    g = (len(d) + 31) // 32
    dd = np.zeros(g * 32, np.int64)
    dd[:len(d)] = np.where(s == 0, d, 0)
    dd = dd.reshape(g, 32)
    it = np.where(dd <= 256, dd, 0).max(axis=1) / 2.0 + np.where(dd > 256, (dd + 127) // 128, 0).sum(axis=1) + 1
    per_warp = np.bincount(np.arange(g) % W, weights=it, minlength=W)
    return int(dd.sum()), per_warp.mean(), per_warp.max()


half = n // 2 // 32 * 32
print(f"N = {n}, adjacency entries = {len(edges)}, warps = {W}")
print("step  infected |  whole network: S-entries  mean  max  |  rank 0 (first half): S-entries  mean  max  |  rank 1: S-entries  mean  max")
tot = np.zeros((3, 2))
steps = 60
for step in range(steps):
    rows = [stats(f.state, 0, n), stats(f.state, 0, half), stats(f.state, half, n)]
    for i, r in enumerate(rows):
        tot[i] += (r[1], r[2])
    if step in (0, 1, 2, 3, 5, 8, 12, 20, 30, 40, 59):
        print(f"{step:4d} {int((f.state == 1).sum()):9d} | " + " | ".join(f"{e / 1e6:8.2f}M {m:7.1f} {x:7.1f}" for e, m, x in rows))
    f.run(1)
print(f"sum over {steps} steps of the per-step (mean, max) warp iterations:")
for name, t in zip(("whole network", "rank 0", "rank 1"), tot):
    print(f"  {name:14s} mean {t[0]:9.0f}   max {t[1]:9.0f}   max/mean {t[1] / t[0]:.2f}")
print(f"  2-rank step = max(rank 0, rank 1) of the slowest warps: {max(tot[1][1], tot[2][1]) / tot[0][1]:.2f} of the single-GPU figure "
      f"(perfect scaling would be 0.50)")
