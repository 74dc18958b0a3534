// schelling.cuh -- Schelling segregation on a Grid (C2): the whole time loop as ONE
// persistent cooperative kernel (cell-binned occupancy, Moore-8 stencil by byte-SWAR adds,
// ordered compaction of the unsatisfied agents, keyed conflict-free matching to empty cells).
//
// Layout in HBM (DESIGN.md "Schelling"):
//   cell_type  int8  [pad | W*H | pad]   -1 empty, else agent type   (env['grid'] of
//              examples/models/schelling_model.py:119-131, packed 4x narrower)
//   cell_agent int32 [W*H]               the binning: which agent sits in the cell
//   E          uint32[e]                 env['empty_cells'] (schelling_model.py:133-139) as
//              cell ids; slot j is refilled with the mover's old cell, so it is maintained
//              in O(movers) per step and never rebuilt
//   type/position/moves per agent        API-visible SoA (schelling_model.py:26-31)
// The pad holds one halo row on each side (0xFF = empty, or the wrapped row when
// Grid(periodic=True), jaxabm/agentpy.py:480) so vertical neighbours are plain offsets.
//
// Per step, inside the kernel (3 grid-wide barriers, no host round trip, no launch):
//   phase 1  every CTA sweeps its contiguous range of 4096-cell tiles: neighbour counts by
//            SWAR byte adds, per-cell satisfaction against a packed threshold table, a 16-bit
//            "unsatisfied" mask per thread (2 B per 16 cells), exact integer partials for the
//            segregation index; publishes its unsatisfied count.
//   phase 2  exclusive prefix over the CTA counts (each CTA folds its predecessors), ordered
//            write of U = unsatisfied cells ascending; CTA 0 writes the step's metrics row.
//   phase 3  mover k < min(u, e): agent in U[piU(k)] -> E[piE(k)], E[piE(k)] <- old cell;
//            piU / piE are keyed Feistel bijections with round keys bits(coll_key, (8,)).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace jxb {

namespace cg = cooperative_groups;

constexpr int kTileCells = 4096;          // cells per tile (256 threads x 16)
constexpr int kCellsPerThread = 16;       // one uint4 of the packed grid

struct SchellingDev {
  signed char* ct;        // points at cell 0 (halo/pad on both sides)
  int* cell_agent;
  unsigned int* U;        // unsatisfied cells of the current step, ascending
  unsigned int* E;        // empty-cell slots
  int* MA;                // agents moved in the last step (for the lazy 'satisfied' column)
  unsigned short* mask16; // unsatisfied mask per 16 cells
  void* blk_part;                 // BlkPart[2][grid]: per-CTA partials (double-buffered by step parity)
  int W, H;               // W rows (x), H columns (y): cell = x*H + y
  long long cells;
  int ntiles;
  int periodic;
  unsigned int n_empty;   // e: constant (agents are conserved)
  unsigned long long need_lut;    // nibble o: least `same` that satisfies an agent with o occupied neighbours
};

__device__ __forceinline__ unsigned int pack_row(unsigned int w) {
  // per byte: bit0 = occupied, bit4 = occupied && type 1    (-1 = 0xFF empty)
  const unsigned int occ = (~w >> 7) & 0x01010101u;
  const unsigned int t1 = w & occ;
  return occ | (t1 << 4);
}

__device__ __forceinline__ unsigned int pack_cell(int v) {
  return v < 0 ? 0u : (1u | ((unsigned)(v & 1) << 4));
}

__device__ __forceinline__ int ldct(const signed char* p) { return (int)__ldcg(p); }

// ---------------------------------------------------------------------------------------
// decision table: index = h | code << 8, h = (#type1 << 4) | #occupied among the 8 neighbours,
// code = 0 type-0 agent, 1 type-1 agent, 3 empty cell.  Entry packs everything the sweep needs
// so that ONE add per cell accumulates all partials:
//   bits  0..12  same * (840 / occupied)      (840 = lcm(1..8): same/occupied as an exact integer)
//   bit   20     agent is unsatisfied
//   bit   26     agent has at least one occupied neighbour
// ---------------------------------------------------------------------------------------
constexpr int kLutSize = 1024;
constexpr unsigned int kLutUnsat = 1u << 20, kLutOcc = 1u << 26;

__device__ __forceinline__ unsigned int lut_entry(unsigned int idx, unsigned long long need_lut) {
  const unsigned int h = idx & 0xFFu, code = idx >> 8;
  const unsigned int o = h & 0xFu, n1 = h >> 4;
  if (code > 1 || o > 8 || n1 > o) return 0u;
  const unsigned int same = code ? n1 : o - n1;
  const unsigned int need = (unsigned int)(need_lut >> (4 * o)) & 0xFu;
  unsigned int e = 0;
  if (o) e = same * (840u / o) | kLutOcc;
  if (same < need) e |= kLutUnsat;
  return e;
}

// neighbour counts of the 16 cells [c0, c0+16) of one lane; a warp covers 512 consecutive cells.
// Horizontal neighbours across lanes come from shuffles, across warps from 3 extra byte loads.
template <bool FAST>
__device__ __forceinline__ void chunk_counts(const SchellingDev& sd, long long c0, bool in_range,
                                             unsigned int (&mid)[4], unsigned int (&Hc)[4]) {
  const int lane = threadIdx.x & 31;
  const int H = sd.H;
  const signed char* ct = sd.ct;
  if (FAST) {
    // H % 16 == 0: the 16 cells of a lane never straddle a row; rows are 16 B aligned
    unsigned int C[4] = {0, 0, 0, 0}, Pm[4] = {0, 0, 0, 0};
    unsigned int left = 0, right = 0;
    bool own_left = true, own_right = true;
    if (in_range) {
      const uint4 u = __ldcg((const uint4*)(ct + c0 - H));
      const uint4 m = __ldcg((const uint4*)(ct + c0));
      const uint4 d = __ldcg((const uint4*)(ct + c0 + H));
      const int col = (int)(c0 % H);
      own_left = lane == 0 || col == 0;
      own_right = lane == 31 || col + kCellsPerThread == H;
      // the cell left of my 16 (lane 0 of the warp, or my cells start a row: wraps if periodic)
      if (own_left && (col != 0 || sd.periodic)) {
        const long long e = col == 0 ? c0 - 1 + H : c0 - 1;
        left = pack_cell(ldct(ct + e - H)) + pack_cell(ldct(ct + e)) + pack_cell(ldct(ct + e + H));
      }
      if (own_right && (col + kCellsPerThread != H || sd.periodic)) {
        const long long e = col + kCellsPerThread == H ? c0 + kCellsPerThread - H : c0 + kCellsPerThread;
        right = pack_cell(ldct(ct + e - H)) + pack_cell(ldct(ct + e)) + pack_cell(ldct(ct + e + H));
      }
      mid[0] = m.x; mid[1] = m.y; mid[2] = m.z; mid[3] = m.w;
      Pm[0] = pack_row(m.x); Pm[1] = pack_row(m.y); Pm[2] = pack_row(m.z); Pm[3] = pack_row(m.w);
      C[0] = pack_row(u.x) + Pm[0] + pack_row(d.x);
      C[1] = pack_row(u.y) + Pm[1] + pack_row(d.y);
      C[2] = pack_row(u.z) + Pm[2] + pack_row(d.z);
      C[3] = pack_row(u.w) + Pm[3] + pack_row(d.w);
    }
    // whole warp, converged: vertical sums of the neighbouring lanes' edge cells
    const unsigned int sl = __shfl_up_sync(0xffffffffu, C[3], 1) >> 24;
    const unsigned int sr = __shfl_down_sync(0xffffffffu, C[0], 1) & 0xFFu;
    if (!own_left) left = sl;
    if (!own_right) right = sr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned int prev = j == 0 ? (left << 24) : C[j - 1];
      const unsigned int next = j == 3 ? right : C[j + 1];
      const unsigned int l = __funnelshift_l(prev, C[j], 8);   // byte i <- byte i-1
      const unsigned int r = __funnelshift_r(C[j], next, 8);   // byte i <- byte i+1
      Hc[j] = l + C[j] + r - Pm[j];
    }
  } else {
    // generic shape: per-cell byte loads with explicit column handling
    if (in_range) {
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        unsigned int mw = 0, hw = 0;
        for (int b = 0; b < 4; ++b) {
          const long long c = c0 + j * 4 + b;
          int v = -1;
          unsigned int h = 0;
          if (c < sd.cells) {
            v = ldct(ct + c);
            const int col = (int)(c % H);
            for (int dy = -1; dy <= 1; ++dy) {
              int cc = col + dy;
              long long shift = dy;
              if (cc < 0) { if (!sd.periodic) continue; shift += H; }
              if (cc >= H) { if (!sd.periodic) continue; shift -= H; }
              h += pack_cell(ldct(ct + c + shift - H)) + pack_cell(ldct(ct + c + shift + H));
              if (dy != 0) h += pack_cell(ldct(ct + c + shift));
            }
          }
          mw |= ((unsigned int)(v & 0xFF)) << (8 * b);
          hw |= h << (8 * b);
        }
        mid[j] = mw;
        Hc[j] = hw;
      }
    }
  }
}

struct __align__(16) BlkPart {
  unsigned int unsat, occ;          // #unsatisfied, #agents with an occupied neighbour
  unsigned long long num;           // sum of same * 840 / occupied
};

template <bool FAST, int MODE>
__global__ void __launch_bounds__(kThreads) schelling_run_kernel(const SchellingDev sd, const ModelDev md, int steps) {
  __shared__ unsigned int s_lut[kLutSize];
  __shared__ unsigned int s_u32[kThreads / 32];
  __shared__ unsigned long long s_u64[kThreads / 32];
  __shared__ unsigned int s_occ[kThreads / 32];
  __shared__ unsigned int s_rk[8];
  __shared__ unsigned int s_prefix, s_total;
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kThreads / 32;
  constexpr int kChunkCells = 32 * kCellsPerThread;     // 512 cells per warp pass
  const int B = gridDim.x, b = blockIdx.x;
  Ctrl* ctrl = md.ctrl;
  const TypeDev& t = md.t[0];
  // contiguous range of 512-cell chunks owned by this CTA (so that U comes out in cell order)
  const long long nchunks = (sd.cells + kChunkCells - 1) / kChunkCells;
  const long long k0 = nchunks * b / B, k1 = nchunks * (b + 1) / B;
  const unsigned int e = sd.n_empty;
  const int step0 = ctrl->step_in_run;           // 0 at launch; read by everyone before anyone writes
  BlkPart* blk_part = (BlkPart*)sd.blk_part;
  for (int i = tid; i < kLutSize; i += kThreads) s_lut[i] = lut_entry(i, sd.need_lut);

  for (int s = 0; s < steps; ++s) {
    // ------------------------------------------------------------------ phase 1: stencil sweep
    if (tid < 8) {
      const uint32_t* kp = md.keys + (size_t)(step0 + s) * (md.n_types + 1) * 2;
      const Key ck = {kp[0], kp[1]};
      s_rk[tid] = bits_elem<MODE>(ck, tid, 8);
    }
    __syncthreads();
    BlkPart* part = blk_part + (size_t)(s & 1) * B;      // double-buffered by step parity: a CTA may
    unsigned int my_unsat = 0, my_occ = 0;                // start step s+1 while others still fold step s
    unsigned long long my_num = 0;
    for (long long ch = k0 + warp; ch < k1; ch += kWarps) {
      const long long c0 = ch * kChunkCells + (long long)lane * kCellsPerThread;
      const bool in_range = c0 < sd.cells;
      unsigned int mid[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
      unsigned int Hc[4] = {0, 0, 0, 0};
      chunk_counts<FAST>(sd, c0, in_range, mid, Hc);
      unsigned int unsat = 0, acc = 0;
      if (in_range) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // code per byte: bit0 = type, bit1 = empty  (0xFF -> 3)
          const unsigned int cw = (mid[j] & 0x01010101u) | ((mid[j] >> 6) & 0x02020202u);
          const unsigned int lo = __byte_perm(Hc[j], cw, 0x5140);   // [h0, code0, h1, code1]
          const unsigned int hi = __byte_perm(Hc[j], cw, 0x7362);   // [h2, code2, h3, code3]
          const unsigned int e0 = s_lut[lo & 0xFFFFu], e1 = s_lut[lo >> 16];
          const unsigned int e2 = s_lut[hi & 0xFFFFu], e3 = s_lut[hi >> 16];
          acc += e0 + e1 + e2 + e3;
          unsat |= ((e0 >> 20) & 1u) << (4 * j) | ((e1 >> 20) & 1u) << (4 * j + 1) |
                   ((e2 >> 20) & 1u) << (4 * j + 2) | ((e3 >> 20) & 1u) << (4 * j + 3);
        }
        sd.mask16[c0 >> 4] = (unsigned short)unsat;
      }
      my_unsat += (acc >> 20) & 0x1Fu;
      my_occ += acc >> 26;
      my_num += acc & 0xFFFFFu;
    }
    {
      const unsigned int a = warp_sum((int)my_unsat), o = warp_sum((int)my_occ);
      unsigned long long n = my_num;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) n += __shfl_xor_sync(0xffffffffu, n, d);
      if (lane == 0) { s_u32[warp] = a; s_occ[warp] = o; s_u64[warp] = n; }
    }
    __syncthreads();
    if (tid == 0) {
      BlkPart p = {0, 0, 0};
      for (int w = 0; w < kWarps; ++w) { p.unsat += s_u32[w]; p.occ += s_occ[w]; p.num += s_u64[w]; }
      part[b] = p;
    }
    grid.sync();

    // ------------------------------------------------------------------ phase 2: ordered U
    {
      unsigned int before = 0, all = 0, occ = 0;
      unsigned long long num = 0;
      for (int i = tid; i < B; i += kThreads) {
        const uint4 raw = __ldcg((const uint4*)(part + i));
        all += raw.x;
        if (i < b) before += raw.x;
        occ += raw.y;
        num += ((unsigned long long)raw.w << 32) | raw.z;
      }
      before = warp_sum((int)before);
      all = warp_sum((int)all);
      occ = warp_sum((int)occ);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) num += __shfl_xor_sync(0xffffffffu, num, d);
      __syncthreads();
      __shared__ unsigned int s_all[kWarps];
      if (lane == 0) { s_u32[warp] = before; s_all[warp] = all; s_occ[warp] = occ; s_u64[warp] = num; }
      __syncthreads();
      if (tid == 0) {
        unsigned int p = 0, a = 0, oc = 0;
        unsigned long long nm = 0;
        for (int w = 0; w < kWarps; ++w) { p += s_u32[w]; a += s_all[w]; oc += s_occ[w]; nm += s_u64[w]; }
        s_prefix = p;
        s_total = a;
        if (b == 0) {
          // metrics row of this step (pre-step configuration; DESIGN.md "Schelling rule")
          const unsigned int u = a, m = u < e ? u : e;
          const long long ts = ctrl->time_step + 1;
          ctrl->total_moves += m;
          ctrl->n_unsat = u;
          ctrl->n_moved = m;
          const double segv = (double)(float)((double)nm / 840.0 / (double)(oc ? oc : 1));
          const double psat = (double)(float)((double)(t.gn - u) / (double)t.gn);
          if ((ts % md.collect_interval) == 0) {
            double* row = md.metrics + (size_t)ctrl->n_recorded * kMaxMetrics;
            row[0] = psat;
            row[1] = segv;
            row[2] = (double)(int)ctrl->total_moves;
            md.record_steps[ctrl->n_recorded] = (int)ts;
            ctrl->n_recorded += 1;
          }
          md.env[0] = segv;
          md.env[1] = psat;
          md.env[2] = (double)(int)ctrl->total_moves;
          ctrl->time_step = ts;
          ctrl->step_in_run += 1;
        }
      }
      __syncthreads();
    }
    const unsigned int u = s_total;
    const unsigned int m = u < e ? u : e;
    if (u > 0) {                   // uniform across the grid
      unsigned int base = s_prefix;
      for (long long ch0 = k0; ch0 < k1; ch0 += kWarps) {
        const long long ch = ch0 + warp;
        const long long c0 = ch * kChunkCells + (long long)lane * kCellsPerThread;
        unsigned int unsat = (ch < k1 && c0 < sd.cells) ? (unsigned int)sd.mask16[c0 >> 4] : 0u;
        const unsigned int cnt = __popc(unsat);
        unsigned int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += v;
        }
        __syncthreads();
        if (lane == 31) s_u32[warp] = inc;
        __syncthreads();
        unsigned int woff = 0, ttot = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          if (w < warp) woff += s_u32[w];
          ttot += s_u32[w];
        }
        unsigned int pu = base + woff + inc - cnt;
        while (unsat) {
          const int q = __ffs(unsat) - 1;
          unsat &= unsat - 1;
          sd.U[pu++] = (unsigned int)(c0 + q);
        }
        base += ttot;
      }
    }
    if (m == 0) continue;          // uniform: nobody moves, the grid is unchanged
    grid.sync();

    // ------------------------------------------------------------------ phase 3: moves
    {
      const Feistel fu = make_feistel(u, s_rk), fe = make_feistel(e, s_rk + 4);
      for (unsigned int k = (unsigned int)b * kThreads + tid; k < m; k += (unsigned int)B * kThreads) {
        const unsigned int src = __ldcg(sd.U + feistel_permute(fu, k));
        const unsigned int j = feistel_permute(fe, k);
        const unsigned int dst = __ldcg(sd.E + j);
        const int a = __ldcg(sd.cell_agent + src);
        const signed char ty = (signed char)ldct(sd.ct + src);
        sd.E[j] = src;
        sd.MA[k] = a;
        sd.ct[dst] = ty;
        sd.ct[src] = (signed char)-1;
        sd.cell_agent[dst] = a;
        sd.cell_agent[src] = -1;
        if (sd.periodic) {
          const long long H = sd.H, cells = sd.cells;
          if (dst < H) sd.ct[dst + cells] = ty;
          if (dst >= cells - H) sd.ct[(long long)dst - cells] = ty;
          if (src < H) sd.ct[src + cells] = (signed char)-1;
          if (src >= cells - H) sd.ct[(long long)src - cells] = (signed char)-1;
        }
        ((int2*)t.f[1])[a] = make_int2((int)(dst / sd.H), (int)(dst % sd.H));
        atomicAdd((int*)t.f[3] + a, 1);   // fire-and-forget L2 reduction: no load to wait for (ncu: 37 % of the mover stalls)
      }
    }
    grid.sync();
  }
}

// rebuild the cell arrays from the per-agent position/type columns
__global__ void grid_clear_kernel(const SchellingDev sd, long long pad) {
  const long long n = sd.cells + 2 * pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    sd.ct[i - pad] = (signed char)-1;
    if (i < sd.cells) sd.cell_agent[i] = -1;
  }
}

__global__ void grid_scatter_kernel(const SchellingDev sd, const int* type, const int2* pos, long long n,
                                    int* err) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int2 p = pos[i];
    if (p.x < 0 || p.x >= sd.W || p.y < 0 || p.y >= sd.H) { atomicExch(err, 1); continue; }
    const long long c = (long long)p.x * sd.H + p.y;
    const int prev = atomicExch(&sd.cell_agent[c], (int)i);
    if (prev != -1) atomicExch(err, 2);       // two agents in one cell
    const signed char ty = (signed char)type[i];
    sd.ct[c] = ty;
    if (sd.periodic) {
      if (c < sd.H) sd.ct[c + sd.cells] = ty;
      if (c >= sd.cells - sd.H) sd.ct[c - sd.cells] = ty;
    }
  }
}

// env['empty_cells'] = ascending empty cells (schelling_model.py:133-139): two-pass ordered
// compaction over 4096-cell tiles (one CTA per tile, 16 cells per thread), run when the grid is
// (re)built from uploaded positions.  Pass 1 counts the empty cells of every tile; pass 2 folds
// the counts of the preceding tiles (exclusive prefix, fixed order) and writes the tile's cells.
__device__ __forceinline__ unsigned int empty_mask16(const SchellingDev& sd, long long c0) {
  unsigned int mask = 0;
#pragma unroll
  for (int q = 0; q < kCellsPerThread; ++q)
    if (c0 + q < sd.cells && sd.ct[c0 + q] < 0) mask |= 1u << q;
  return mask;
}

__global__ void __launch_bounds__(kThreads) empty_count_kernel(const SchellingDev sd, unsigned int* tile_count) {
  __shared__ unsigned int s_w[kThreads / 32];
  const int tid = threadIdx.x;
  const long long c0 = (long long)blockIdx.x * kTileCells + (long long)tid * kCellsPerThread;
  const unsigned int cnt = warp_sum((int)__popc(empty_mask16(sd, c0)));
  if ((tid & 31) == 0) s_w[tid >> 5] = cnt;
  __syncthreads();
  if (tid == 0) {
    unsigned int tot = 0;
    for (int w = 0; w < kThreads / 32; ++w) tot += s_w[w];
    tile_count[blockIdx.x] = tot;
  }
}

__global__ void __launch_bounds__(kThreads) empty_write_kernel(const SchellingDev sd, const unsigned int* tile_count,
                                                               unsigned int* count_out) {
  __shared__ unsigned int s_w[kThreads / 32];
  __shared__ unsigned int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // exclusive prefix of the preceding tiles' counts
  unsigned int before = 0;
  for (int i = tid; i < (int)blockIdx.x; i += kThreads) before += tile_count[i];
  before = warp_sum((int)before);
  if (lane == 0) s_w[warp] = before;
  __syncthreads();
  if (tid == 0) {
    unsigned int tot = 0;
    for (int w = 0; w < kThreads / 32; ++w) tot += s_w[w];
    s_base = tot;
  }
  __syncthreads();
  const long long c0 = (long long)blockIdx.x * kTileCells + (long long)tid * kCellsPerThread;
  unsigned int mask = empty_mask16(sd, c0);
  const unsigned int cnt = __popc(mask);
  unsigned int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  __syncthreads();
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  unsigned int woff = 0, ttot = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    if (w < warp) woff += s_w[w];
    ttot += s_w[w];
  }
  unsigned int pos = s_base + woff + inc - cnt;
  while (mask) {
    const int q = __ffs(mask) - 1;
    mask &= mask - 1;
    if (pos < sd.n_empty) sd.E[pos] = (unsigned int)(c0 + q);   // more only if two agents share a cell
    ++pos;
  }
  if (blockIdx.x == gridDim.x - 1 && tid == 0) *count_out = s_base + ttot;
}

__global__ void grid_export_kernel(const SchellingDev sd, int* out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < sd.cells;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = (int)sd.ct[i];
}

// 'satisfied' column (schelling_model.py:29) of the last step, materialised on demand: everyone
// is satisfied except the agents moved in that step (MA) and the unsatisfied agents that stayed
// (still sitting in their U cell)
__global__ void satisfied_export_kernel(const SchellingDev sd, const Ctrl* ctrl, unsigned char* sat) {
  const unsigned int u = ctrl->n_unsat, m = ctrl->n_moved;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < u;
       i += (long long)gridDim.x * blockDim.x) {
    if (i < m) sat[sd.MA[i]] = 0;
    const int a = sd.cell_agent[sd.U[i]];
    if (a >= 0) sat[a] = 0;
  }
}

}  // namespace jxb
