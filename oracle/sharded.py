"""CPU restatement of the MULTI-GPU decompositions (TEST INFRASTRUCTURE ONLY, like the rest of ``oracle/``).

``csrc/grid_shard.cuh`` splits ONE Schelling grid into row bands and ``csrc/sir.cuh`` ONE SIR network into
node ranges; the reference has no multi-device path (``jaxabm/agentpy.py:480-527`` keeps one
``env['grid']``, ``:545-582`` one ``env['network_edges']``), so these functions restate the *decomposed*
algorithms rank by rank in NumPy -- every rank only reads what the device kernels let it read -- and
``tests/test_oracle_sharded.py`` checks that they reproduce the plain oracle (``oracle/rules.py``) exactly.
That pins the design decisions the kernels rely on without a GPU:

  * a band only needs rows [X0-1, X1] of the grid; they stay coherent when every mover's record goes to the
    owner of its target row AND to the ranks that keep the target row -- or the source row -- as a halo
    (including the wrapped rows of a periodic grid): no halo exchange besides the movers themselves;
  * the ordered per-band unsatisfied lists, concatenated in rank order, are the global list ``U``, so a rank
    can walk ITS movers alone: entry j of its segment is mover k = piU^-1(j);
  * ``empty_cells`` can be partitioned by slot range: a slot is matched to exactly one mover per step, so its
    owner can serve the movers' requests in any order and forward each mover to the owner of the target row;
  * the agent id and its move count travel with the cell (the record), so no rank needs a per-agent column
    during the run; columns combine as position = max, satisfied = min, moves = sum over the ranks' views;
  * a node range of the SIR network only needs its own rows + the global infected bitmap.
"""
from __future__ import annotations

import numpy as np

from . import jaxlike as jl
from .rules import moore_counts

i32, f32 = np.int32, np.float32
STALE = -7          # marks grid rows a rank must never read (outside its band + halo)


def _bounds(n, rank, world):
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _BandRank:
    """One rank's state: its own copy of the grid (only rows [X0-1, X1] are meaningful), the payload
    (agent, moves) of the cells of its rows, and ITS range of the empty-cell slots."""

    def __init__(self, rank, world, W, H, periodic, types, positions, threshold):
        self.rank, self.world, self.W, self.H, self.periodic = rank, world, W, H, periodic
        self.X0, self.X1 = _bounds(W, rank, world)
        self.xb = [_bounds(W, q, world)[0] for q in range(world)] + [W]
        self.thr = f32(threshold)
        self.n = len(types)
        grid = -np.ones((W, H), dtype=i32)
        grid[positions[:, 0], positions[:, 1]] = types
        empty = np.flatnonzero(grid.reshape(-1) < 0).astype(np.int64)      # slot order = ascending at start
        self.e = len(empty)
        self.eper = max(1, -(-self.e // world))
        self.E = empty[rank * self.eper:(rank + 1) * self.eper].copy()     # my slots only
        self.grid = np.full((W, H), STALE, dtype=i32)
        for x in self._kept_rows():
            self.grid[x] = grid[x]
        self.cell_agent = -np.ones((W, H), dtype=np.int64)          # rows [X0, X1) only
        self.cell_moves = np.zeros((W, H), dtype=np.int64)
        own = (positions[:, 0] >= self.X0) & (positions[:, 0] < self.X1)
        self.cell_agent[positions[own, 0], positions[own, 1]] = np.flatnonzero(own)
        self.inbox = []                                             # records received this step
        self.requests = []                                          # movers that take one of my slots this step
        self.last_unsat = np.zeros(0, dtype=np.int64)

    def _kept_rows(self):
        rows = set(range(max(self.X0 - 1, 0), min(self.X1 + 1, self.W)))
        if self.periodic:
            rows.add((self.X0 - 1) % self.W)
            rows.add(self.X1 % self.W)
        return rows

    def _owner(self, x):
        return max(q for q in range(self.world) if x >= self.xb[q])

    def _halo_ranks(self, x, own):
        """The ranks other than ``own`` that keep row x as a halo row: the owners of the rows next to it."""
        out = []
        for xn in (x - 1, x + 1):
            if self.periodic:
                xn %= self.W
            if 0 <= xn < self.W:
                q = self._owner(xn)
                if q != own and q not in out:
                    out.append(q)
        return out

    def sweep(self):
        """Unsatisfied agents of my rows from rows [X0-1, X1] of MY copy -> ordered cells + integer partials."""
        X0, X1, W = self.X0, self.X1, self.W
        rows = [(x % W) if self.periodic else x for x in range(X0 - 1, X1 + 1)]
        pad = -np.ones((1, self.H), dtype=i32)
        window = np.concatenate([self.grid[[x]] if 0 <= x < W else pad for x in rows], axis=0)
        assert not np.any(window == STALE), "a band read a row outside its band + halo"
        if self.periodic:
            # wrap the columns only; the two outer rows of the window are the halo rows
            occ = (window >= 0).astype(i32); t0 = (window == 0).astype(i32); t1 = (window == 1).astype(i32)

            def nsum(a):
                tot = np.zeros_like(a[1:-1])
                for dx in (0, 1, 2):
                    for dy in (-1, 0, 1):
                        if dx != 1 or dy:
                            tot += np.roll(a[dx:dx + X1 - X0], dy, axis=1)
                return tot
            o, n0, n1 = nsum(occ), nsum(t0), nsum(t1)
        else:
            o, n0, n1 = (a[1:-1] for a in moore_counts(window, False))
        band = window[1:-1]
        same = np.where(band == 0, n0, n1)
        with np.errstate(divide="ignore", invalid="ignore"):
            frac = (same.astype(f32) / o.astype(f32)).astype(f32)
        agent = band >= 0
        unsat = agent & (o > 0) & ~(frac >= self.thr)
        xs, ys = np.nonzero(unsat)                                  # row-major = ascending cell id
        self.U = (xs + X0).astype(np.int64) * self.H + ys           # my segment of the global list
        self.last_unsat = self.cell_agent[xs + X0, ys].copy()
        sel = agent & (o > 0)
        num = int(np.sum(same[sel].astype(np.int64) * (840 // np.maximum(o[sel], 1))))
        return len(self.U), int(sel.sum()), num

    def moveout(self, ranks, prefix, u, m, rk):
        """Walk MY movers: entry j of my segment is mover k with piU(k) = j; it takes slot piE(k).  The source cell
        is cleared here; the mover leaves as a request to the rank that holds the slot."""
        k = np.arange(m, dtype=np.uint32)
        j = jl.feistel_permute(k, u, rk[0:4]).astype(np.int64)
        mine = (j >= prefix[self.rank]) & (j < prefix[self.rank + 1])
        slots = jl.feistel_permute(k[mine], self.e, rk[4:8]).astype(np.int64)
        H = self.H
        for jj, sl in zip(j[mine] - prefix[self.rank], slots):
            s_ = int(self.U[jj])
            xs = s_ // H
            a, mv, ty = self.cell_agent[xs, s_ % H], self.cell_moves[xs, s_ % H], self.grid[xs, s_ % H]
            assert a >= 0 and ty >= 0
            self._flip(s_, -1)
            self.cell_agent[xs, s_ % H] = -1
            self.cell_moves[xs, s_ % H] = 0
            for p in self._halo_ranks(xs, self.rank):
                ranks[p].inbox.append((s_, -1, 0, -1))
            q, off = divmod(int(sl), self.eper)
            ranks[q].requests.append((off, s_, int(a), int(mv), int(ty)))

    def forward(self, ranks):
        """The slot owner's part: slot -> target cell, slot <- source cell; the mover goes on to the owner of the
        target row, plane-only copies to the ranks that keep that row as a halo."""
        for off, s_, a, mv, ty in self.requests:
            d_ = int(self.E[off])
            self.E[off] = s_
            xd = d_ // self.H
            p = self._owner(xd)
            rec = (d_, a, mv + 1, ty)
            ranks[p].inbox.append(rec)
            for ph in self._halo_ranks(xd, p):
                ranks[ph].inbox.append(rec)
        self.requests = []

    def _flip(self, c, value):
        x = c // self.H
        if x in self._kept_rows():
            self.grid[x, c % self.H] = value

    def apply(self):
        for c, a, mv, ty in self.inbox:
            self._flip(c, ty)
            x = c // self.H
            if ty >= 0 and self.X0 <= x < self.X1:
                self.cell_agent[x, c % self.H] = a
                self.cell_moves[x, c % self.H] = mv
        self.inbox = []

    def export(self):
        """My view of the per-agent columns after the last step (what jxb_model_download returns)."""
        pos = -np.ones((self.n, 2), dtype=i32)
        moves = np.zeros(self.n, dtype=i32)
        xs, ys = np.nonzero(self.cell_agent[self.X0:self.X1] >= 0)
        ag = self.cell_agent[xs + self.X0, ys]
        pos[ag] = np.stack([xs + self.X0, ys], axis=1)
        moves[ag] = self.cell_moves[xs + self.X0, ys]
        sat = np.ones(self.n, dtype=bool)
        sat[self.last_unsat] = False
        return pos, sat, moves


def schelling_bands_run(grid_size, types, positions, world, steps, seed_key, mode, threshold=0.5, periodic=False):
    """Run ``steps`` steps of the band-decomposed Schelling model on ``world`` ranks -> (state dict, metric rows,
    empty_cells slots).  ``seed_key`` is the model's root key (``PRNGKey(config.seed)``); the key schedule is
    ``jaxabm/model.py:129-130,156,164,183`` for one collection with an env function."""
    W = H = grid_size
    n = len(types)
    ranks = [_BandRank(r, world, W, H, periodic, types, positions, threshold) for r in range(world)]
    e = ranks[0].e
    rng = jl.split(seed_key, 2, mode)[0]                          # initialize(): keys = split(rng, C + 1), rng = keys[0]
    rows, total_moves = [], 0
    for _ in range(steps):
        rng, step_key = jl.split(rng, 2, mode)
        step_key, coll_key = jl.split(step_key, 2, mode)
        rk = jl.random_bits(coll_key, (8,), mode)
        counts = [rk_.sweep() for rk_ in ranks]                   # the bands' counts, folded in rank order everywhere
        u = sum(c[0] for c in counts); occ = sum(c[1] for c in counts); num = sum(c[2] for c in counts)
        prefix = np.concatenate([[0], np.cumsum([c[0] for c in counts])])
        m = min(u, e)
        if m > 0:
            for r in ranks:
                r.moveout(ranks, prefix, u, m, rk)
            for r in ranks:
                r.forward(ranks)
            for r in ranks:
                r.apply()
        total_moves += m
        rows.append((f32((n - u) / n), f32(num / 840.0 / max(occ, 1)), total_moves))
    views = [r.export() for r in ranks]
    pos = np.max(np.stack([v[0] for v in views]), axis=0)
    sat = np.min(np.stack([v[1] for v in views]), axis=0)
    moves = np.sum(np.stack([v[2] for v in views]), axis=0).astype(i32)
    E = np.concatenate([r.E for r in ranks])
    return {"type": types, "position": pos, "satisfied": sat, "moves": moves}, rows, E


def sir_node_ranges_run(n, edges, cuts, steps, seed_key, mode, beta=0.05, gamma=0.1, initial_infected=0.01,
                        local_edges=None):
    """SIR (``oracle/rules.py::SIRAgent``) computed range by range: rank r holds the CSR of the edges whose
    SOURCE lies in ``[cuts[r], cuts[r+1])`` (sources re-based to local rows, targets global ids -- the filter of
    ``jaxabm_b200/model.py::_push_env``), its slice of ``state``, the draws of its agents by GLOBAL index
    (``split(key, N)[lo:hi]``, ``jaxabm/agent.py:115,156``) and a copy of the global infected bitmap that is
    reassembled from the ranks' slices after every step.  ``local_edges`` lets a test substitute the product's
    own edge filter (``jaxabm_b200.sharding.local_edges``).  -> (final state, [(S, I, R)] per step)."""
    from .rules import SIR_KCAP, edges_to_csr, sir_escape_table
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    q = sir_escape_table(beta)
    world = len(cuts) - 1
    keys = jl.split(seed_key, 2, mode)                             # initialize(): rng = keys[0], collection <- keys[1]
    rng = keys[0]
    agent_keys = jl.split(keys[1], n, mode)
    local = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        mine = edges[(edges[:, 0] >= lo) & (edges[:, 0] < hi)].copy() if local_edges is None else \
            np.asarray(local_edges(edges, lo, hi), dtype=np.int64)
        if local_edges is None:
            mine[:, 0] -= lo
        row_ptr, col = edges_to_csr(hi - lo, mine)
        u0 = jl.uniform_scalar_batched(agent_keys[lo:hi], mode=mode)
        local.append({"lo": lo, "hi": hi, "row_ptr": row_ptr, "col": col.astype(np.int64),
                      "state": (u0 < f32(initial_infected)).astype(i32)})
    bitmap = np.concatenate([r["state"] == 1 for r in local])     # every rank's copy after the initial sync
    rows = []
    for _ in range(steps):
        rng, step_key = jl.split(rng, 2, mode)
        _, coll_key = jl.split(step_key, 2, mode)
        ks = jl.split(coll_key, n, mode)
        new_slices, counts = [], np.zeros(3, dtype=np.int64)
        for r in local:
            st = r["state"]
            inf = bitmap[r["col"]].astype(np.int64)                # gathers hit the GLOBAL bitmap
            csum = np.concatenate([[0], np.cumsum(inf)])
            k = csum[r["row_ptr"][1:]] - csum[r["row_ptr"][:-1]]
            u = jl.uniform_scalar_batched(ks[r["lo"]:r["hi"]], mode=mode)
            p_inf = (f32(1.0) - q[np.minimum(k, SIR_KCAP)]).astype(f32)
            new = st.copy()
            new[(st == 0) & (u < p_inf)] = 1
            new[(st == 1) & (u < f32(gamma))] = 2
            r["state"] = new
            new_slices.append(new == 1)
            counts += np.array([np.sum(new == 0), np.sum(new == 1), np.sum(new == 2)])
        bitmap = np.concatenate(new_slices)                        # the pull kernels' remote word stores
        rows.append(tuple(int(c) for c in counts))
    return np.concatenate([r["state"] for r in local]), rows
