// schelling.cuh -- Schelling segregation on a Grid (C2): cell-binned occupancy, Moore-8
// stencil, single-pass stream compaction of movers / empty cells, keyed matching.
//
// Layout in HBM (DESIGN.md "Schelling"):
//   cell_type  int8  [pad | W*H | pad]   -1 empty, else agent type   (env['grid'] of
//              examples/models/schelling_model.py:119-131, packed 4x narrower)
//   cell_agent int32 [W*H]               the binning: which agent sits in the cell
//   type/position/moves per agent        API-visible SoA (schelling_model.py:26-31)
// The pad holds one halo row on each side (0xFF = empty, or the wrapped row when
// Grid(periodic=True), jaxabm/agentpy.py:480) so vertical neighbours are plain offsets.
//
// Step = 2 launches:
//   stencil_compact_kernel : per 4096-cell tile, neighbour counts by byte-SWAR adds,
//       satisfied / empty flags, block scan + decoupled look-back across tiles, and the
//       ordered lists U (unsatisfied cells, ascending cell id), UA (their agents) and
//       E (empty cells, ascending) written straight from registers -- no flag array.
//   move_kernel : mover k < min(u,e): U[piU(k)] -> E[piE(k)] with two keyed Feistel
//       bijections (round keys = bits(coll_key, (8,))); conflict-free by construction;
//       last CTA folds the per-tile partials and writes the step's metrics row.
#pragma once
#include "common.cuh"

namespace jxb {

constexpr int kTileCells = 4096;          // cells per CTA tile
constexpr int kCellsPerThread = 16;       // one uint4 of the packed grid

struct SchellingDev {
  signed char* ct;        // points at cell 0 (halo/pad on both sides)
  int* cell_agent;
  unsigned int* U;        // unsatisfied cells, ascending
  int* UA;                // agent sitting in U[k]
  unsigned int* E;        // empty cells, ascending
  unsigned long long* tile_desc;
  double* tile_seg_sum;   // per-tile sum of same/occupied
  int* tile_seg_cnt;
  int W, H;               // W rows (x), H columns (y): cell = x*H + y
  long long cells;
  int ntiles;
  int periodic;
  unsigned int sat_lut[10];     // bit s of sat_lut[o]: satisfied with s same of o occupied
  const float* ratio_lut;       // [10*16] same/occupied in float32 (0 where occ == 0)
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

constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagPre = 2ull << 62;
constexpr unsigned long long kCntMask = (1ull << 31) - 1;

template <bool FAST>
__global__ void __launch_bounds__(kThreads) stencil_compact_kernel(const SchellingDev sd, Ctrl* ctrl) {
  __shared__ unsigned int sC[kTileCells / 4 + 2];   // vertical sums, packed, +1 word each side
  __shared__ unsigned int s_warp[kThreads / 32];
  __shared__ unsigned int s_tile, s_base_u, s_base_e;
  __shared__ double s_seg[kThreads / 32];
  __shared__ int s_cnt[kThreads / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(&ctrl->tile_ticket, 1u);
  __syncthreads();
  const unsigned int tile = s_tile;
  const long long c0 = (long long)tile * kTileCells + (long long)tid * kCellsPerThread;
  const int H = sd.H;
  const signed char* ct = sd.ct;

  unsigned int mid[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
  unsigned int Hc[4];                     // per byte: low nibble #occupied, high nibble #type1
  bool in_range = c0 < sd.cells;

  if (FAST) {
    // H % 16 == 0: the 16 cells of a thread never straddle a row; rows are 16B aligned
    unsigned int C[4] = {0, 0, 0, 0}, Pm[4] = {0, 0, 0, 0};
    if (in_range) {
      const uint4 u = *(const uint4*)(ct + c0 - H);
      const uint4 m = *(const uint4*)(ct + c0);
      const uint4 d = *(const uint4*)(ct + c0 + H);
      mid[0] = m.x; mid[1] = m.y; mid[2] = m.z; mid[3] = m.w;
      Pm[0] = pack_row(m.x); Pm[1] = pack_row(m.y); Pm[2] = pack_row(m.z); Pm[3] = pack_row(m.w);
      C[0] = pack_row(u.x) + Pm[0] + pack_row(d.x);
      C[1] = pack_row(u.y) + Pm[1] + pack_row(d.y);
      C[2] = pack_row(u.z) + Pm[2] + pack_row(d.z);
      C[3] = pack_row(u.w) + Pm[3] + pack_row(d.w);
    }
    unsigned int* myC = sC + 1 + tid * 4;
    myC[0] = C[0]; myC[1] = C[1]; myC[2] = C[2]; myC[3] = C[3];
    // tile-edge words: the cell just left of the tile and just right of it
    if (tid == 0) {
      const long long e = (long long)tile * kTileCells - 1;
      sC[0] = (pack_cell(ct[e - H]) + pack_cell(ct[e]) + pack_cell(ct[e + H])) << 24;
    }
    if (tid == kThreads - 1) {
      const long long e = (long long)(tile + 1) * kTileCells;
      unsigned int v = 0;
      if (e < sd.cells + H) v = pack_cell(ct[e - H]) + pack_cell(ct[e]) + pack_cell(ct[e + H]);
      sC[kTileCells / 4 + 1] = v;
    }
    __syncthreads();
    if (in_range) {
      unsigned int left = myC[-1] >> 24;        // vertical sum of the cell left of my 16
      unsigned int right = myC[4] & 0xFFu;      // and right of them
      const int col = (int)(c0 % H);
      if (col == 0) {
        left = 0;
        if (sd.periodic) {
          const long long e = c0 - 1 + H;       // (row, H-1) and its vertical neighbours
          left = pack_cell(ct[e - H]) + pack_cell(ct[e]) + pack_cell(ct[e + H]);
        }
      }
      if (col + kCellsPerThread == H) {
        right = 0;
        if (sd.periodic) {
          const long long e = c0 + kCellsPerThread - H;   // (row, 0)
          right = pack_cell(ct[e - H]) + pack_cell(ct[e]) + pack_cell(ct[e + H]);
        }
      }
      // horizontal 3-sum per byte, minus the centre cell
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned int prev = j == 0 ? (left << 24) : C[j - 1];
        const unsigned int next = j == 3 ? right : C[j + 1];
        const unsigned int l = __funnelshift_l(prev, C[j], 8);   // byte i <- byte i-1
        const unsigned int r = __funnelshift_r(C[j], next, 8);   // byte i <- byte i+1
        Hc[j] = l + C[j] + r - Pm[j];
      }
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
            v = ct[c];
            const int col = (int)(c % H);
            for (int dy = -1; dy <= 1; ++dy) {
              int cc = col + dy;
              long long shift = dy;
              if (cc < 0) { if (!sd.periodic) continue; shift += H; }
              if (cc >= H) { if (!sd.periodic) continue; shift -= H; }
              h += pack_cell(ct[c + shift - H]) + pack_cell(ct[c + shift + H]);
              if (dy != 0) h += pack_cell(ct[c + shift]);
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

  // ---- per-cell decisions -----------------------------------------------------------
  unsigned int unsat = 0, empty = 0;      // 16-bit masks over my cells
  float seg = 0.f;
  int segc = 0;
  if (in_range) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int idx = j * 4 + b;
        if (c0 + idx >= sd.cells) continue;
        const unsigned int v = (mid[j] >> (8 * b)) & 0xFFu;
        if (v == 0xFFu) {
          empty |= 1u << idx;
          continue;
        }
        const unsigned int h = (Hc[j] >> (8 * b)) & 0xFFu;
        const unsigned int o = h & 0xFu, n1 = h >> 4;
        const unsigned int same = v ? n1 : o - n1;
        const unsigned int sat = (sd.sat_lut[o] >> same) & 1u;
        if (!sat) unsat |= 1u << idx;
        if (o) {
          seg += __ldg(sd.ratio_lut + o * 16 + same);
          segc += 1;
        }
      }
    }
  }

  // ---- block scan of (unsat | empty << 16) counts ------------------------------------
  const unsigned int cnt = __popc(unsat) | (__popc(empty) << 16);
  unsigned int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) s_warp[warp] = inc;
  {
    double sv = warp_sum((double)seg);
    int sc = warp_sum(segc);
    if (lane == 0) { s_seg[warp] = sv; s_cnt[warp] = sc; }
  }
  __syncthreads();
  unsigned int warp_off = 0, block_tot = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    if (w < warp) warp_off += s_warp[w];
    block_tot += s_warp[w];
  }
  const unsigned int excl = warp_off + inc - cnt;

  // ---- decoupled look-back across tiles (warp 0) --------------------------------------
  if (warp == 0) {
    const unsigned long long agg =
        ((unsigned long long)(block_tot & 0xFFFFu) << 31) | (unsigned long long)(block_tot >> 16);
    unsigned long long prefix = 0;
    volatile unsigned long long* desc = sd.tile_desc;
    if (tile == 0) {
      if (lane == 0) desc[0] = kFlagPre | agg;
    } else {
      if (lane == 0) desc[tile] = kFlagAgg | agg;
      long long pred = (long long)tile - 1 - lane;
      while (true) {
        unsigned long long d = 0;
        if (pred >= 0) {
          do { d = desc[pred]; } while ((d >> 62) == 0);
        } else {
          d = kFlagPre;   // before tile 0: zero prefix
        }
        const unsigned int is_pre = __ballot_sync(0xffffffffu, (d >> 62) == 2);
        unsigned long long contrib = d & ((1ull << 62) - 1);
        if (is_pre) {
          const int first = __ffs(is_pre) - 1;       // nearest predecessor holding a full prefix
          if (lane > first) contrib = 0;
        }
        // sum the two 31-bit fields separately (no cross-field carry: totals < 2^31)
        unsigned long long a = contrib >> 31, b = contrib & kCntMask;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        prefix += (a << 31) | b;
        if (is_pre) break;
        pred -= 32;
      }
      if (lane == 0) {
        const unsigned long long incl =
            (((prefix >> 31) + (agg >> 31)) << 31) | ((prefix & kCntMask) + (agg & kCntMask));
        __threadfence();
        desc[tile] = kFlagPre | incl;
      }
    }
    if (lane == 0) {
      s_base_u = (unsigned int)(prefix >> 31);
      s_base_e = (unsigned int)(prefix & kCntMask);
      double sv = 0; int sc = 0;
      for (int w = 0; w < kThreads / 32; ++w) { sv += s_seg[w]; sc += s_cnt[w]; }
      sd.tile_seg_sum[tile] = sv;
      sd.tile_seg_cnt[tile] = sc;
      if ((int)tile == sd.ntiles - 1) {
        ctrl->n_unsat = s_base_u + (block_tot & 0xFFFFu);
        ctrl->n_empty = s_base_e + (block_tot >> 16);
      }
    }
  }
  __syncthreads();

  // ---- ordered writes -------------------------------------------------------------------
  unsigned int pu = s_base_u + (excl & 0xFFFFu);
  unsigned int pe = s_base_e + (excl >> 16);
  while (unsat) {
    const int b = __ffs(unsat) - 1;
    unsat &= unsat - 1;
    const unsigned int c = (unsigned int)(c0 + b);
    sd.U[pu] = c;
    sd.UA[pu] = sd.cell_agent[c];
    ++pu;
  }
  while (empty) {
    const int b = __ffs(empty) - 1;
    empty &= empty - 1;
    sd.E[pe++] = (unsigned int)(c0 + b);
  }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads) move_kernel(const SchellingDev sd, const ModelDev md) {
  __shared__ unsigned int s_rk[8];
  __shared__ int s_last;
  __shared__ double s_d[kThreads / 32];
  __shared__ long long s_c[kThreads / 32];
  Ctrl* ctrl = md.ctrl;
  const unsigned int u = ctrl->n_unsat, e = ctrl->n_empty;
  const unsigned int m = u < e ? u : e;
  const TypeDev& t = md.t[0];
  if (threadIdx.x < 8) {
    const int step = ctrl->step_in_run;
    const uint32_t* kp = md.keys + (size_t)step * (md.n_types + 1) * 2;
    Key ck = {kp[0], kp[1]};
    s_rk[threadIdx.x] = bits_elem<MODE>(ck, threadIdx.x, 8);
  }
  __syncthreads();
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < sd.ntiles) sd.tile_desc[k] = 0;     // re-arm the look-back for the next step
  if (k < m) {
    const Feistel fu = make_feistel(u, s_rk), fe = make_feistel(e, s_rk + 4);
    const unsigned int ku = feistel_permute(fu, (unsigned int)k);
    const unsigned int src = sd.U[ku];
    const int a = sd.UA[ku];
    const unsigned int dst = sd.E[feistel_permute(fe, (unsigned int)k)];
    const signed char ty = sd.ct[src];
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
    ((int*)t.f[3])[a] += 1;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&ctrl->ticket2, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  // ---- tail: fold per-tile partials in tile order, write the metrics row ----------------
  double sv = 0;
  long long sc = 0;
  for (int i = threadIdx.x; i < sd.ntiles; i += blockDim.x) {
    sv += __ldcg(sd.tile_seg_sum + i);
    sc += __ldcg(sd.tile_seg_cnt + i);
  }
  sv = warp_sum(sv);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
  if ((threadIdx.x & 31) == 0) { s_d[threadIdx.x >> 5] = sv; s_c[threadIdx.x >> 5] = sc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    sv = 0; sc = 0;
    for (int w = 0; w < kThreads / 32; ++w) { sv += s_d[w]; sc += s_c[w]; }
    ctrl->ticket2 = 0;
    ctrl->tile_ticket = 0;
    ctrl->total_moves += m;
    ctrl->n_satisfied = t.gn - u;
    ctrl->seg_sum = sv;
    ctrl->seg_cnt = sc;
    const long long ts = ctrl->time_step + 1;
    if ((ts % md.collect_interval) == 0) {
      double* row = md.metrics + (size_t)ctrl->n_recorded * kMaxMetrics;
      row[0] = (double)(float)((double)(t.gn - u) / (double)t.gn);          // percent_satisfied
      row[1] = (double)(float)(sv / (double)(sc > 0 ? sc : 1));             // segregation_index
      row[2] = (double)(int)ctrl->total_moves;                              // total_moves (int32)
      md.record_steps[ctrl->n_recorded] = (int)ts;
      ctrl->n_recorded += 1;
    }
    md.env[0] = (double)(float)(sv / (double)(sc > 0 ? sc : 1));
    md.env[1] = (double)(float)((double)(t.gn - u) / (double)t.gn);
    md.env[2] = (double)(int)ctrl->total_moves;
    ctrl->time_step = ts;
    ctrl->step_in_run += 1;
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

__global__ void grid_export_kernel(const SchellingDev sd, int* out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < sd.cells;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = (int)sd.ct[i];
}

// 'satisfied' column (schelling_model.py:29) of the last step, materialised on demand:
// everyone is satisfied except the agents listed in UA[0..n_unsat)
__global__ void satisfied_export_kernel(const SchellingDev sd, const Ctrl* ctrl, unsigned char* sat) {
  const unsigned int u = ctrl->n_unsat;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < u;
       i += (long long)gridDim.x * blockDim.x)
    sat[sd.UA[i]] = 0;
}

}  // namespace jxb
