// schelling_bits.cuh -- bit-sliced variant of the persistent Schelling kernel (C2) for grids whose
// row length is a multiple of 1024 cells (the bench shape 4096 x 4096 among them).
//
// The grid is held as two bit planes, 1 bit per cell, 32 cells per word, row-major with one halo
// row on each side (zero, or the wrapped row when Grid(periodic=True)):
//   occ : cell is occupied            t1 : cell is occupied by a type-1 agent
// (env['grid'] of examples/models/schelling_model.py:119-131 packed 32x narrower; the int8 grid of
// schelling.cuh is rebuilt from the planes when the host asks for it).
//
// A lane owns one word column and walks down the rows of its CTA's band; a warp covers 32
// consecutive words (1024 cells) of a row.  For each row it forms, per plane, the horizontal
// 3-cell sums (2 bit-sliced bits) of the rows above/below and the 2-cell sum of the own row --
// each row's sums are computed once and reused for three output rows -- adds them with a
// carry-save network into the 4-bit neighbour counts o (occupied) and n1 (type 1), all 32 cells of
// the word at once.  same = type ? n1 : o - n1 (bit-sliced subtract + mux); the threshold table
// need[o] of the rule (float32 same/occupied >= threshold, evaluated on the host for o = 1..8) is
// applied as a bit-sliced constant select + 4-bit comparator; the segregation numerator
// sum(same * 840 / o) is accumulated as popcounts of (o == k) & bit_b(same).  ~6 integer
// instructions per cell instead of ~25 for the byte/LUT sweep; results are bit-identical.
//
// Phase 2 (ordered compaction of the unsatisfied cells) and phase 3 (keyed Feistel matching of movers to empty-cell
// slots, the same matching as in schelling.cuh) are CTA-local here: every CTA compacts the unsatisfied cells of ITS
// rows into its segment of U and walks that segment in cell order, asking the inverse permutation which mover index
// each entry is.  The agent id and its move count travel with the cell (cell_am); 'position' / 'moves' are derived
// from it when they are read.  Plane bits are flipped by atomicAnd/Or.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "schelling.cuh"

#ifndef JXB_SCH_MV
#define JXB_SCH_MV 4      // list entries per thread per iteration of the mover loop (measured: profiles/r02_schelling_mv.txt)
#endif
#ifndef JXB_SCH_MINB
#define JXB_SCH_MINB 2    // resident CTAs per SM the register allocation aims at (128 registers at 2)
#endif

namespace jxb {

struct SchellingBitsDev {
  unsigned int* occ;           // word 0 of row 0 (a halo row of wpr words before and after)
  unsigned int* t1;
  unsigned int* umask;         // [cells / 32] unsatisfied agents of the step
  int2* cell_am;               // persistent kernel only: per cell (agent id or -1, that agent's moves) -- the payload that
                               // travels with a mover; 'position' / 'moves' columns are derived from it on demand
  int wpr;                     // words per row (H / 32)
  int spr;                     // 1024-cell strips per row (wpr / 32): 1, 2, 4 or 8
  unsigned int need_sel[9][4]; // need_sel[o][b] = all-ones iff bit b of need[o] is set (o = 1..8)
};

// horizontal sums of one row of one plane for the 32 cells of a lane's word
struct RowSums {
  unsigned int c;        // the word itself
  unsigned int s0, s1;   // left + centre + right   (2 bit-sliced bits)
  unsigned int m0, m1;   // left + right
};

__device__ __forceinline__ unsigned int maj3(unsigned int a, unsigned int b, unsigned int c) {
  return (a & b) | (a & c) | (b & c);
}

// row x of a plane (x may be -1 or W: the halo rows).  Every lane loads its own word and the words
// left and right of it (indices jl / jr precomputed per lane, wrapped when periodic; lmask / rmask
// zero the neighbour at a non-periodic row end) -- branch-free, no shuffles; the three loads of a
// lane hit the same or the adjacent L2 sector.  fetch_row only issues the loads (so that the next
// row is in flight while the current one is evaluated), finish_row forms the horizontal sums.
struct RawRow {
  unsigned int c, l, r;
};

struct LaneCols {
  int j, jl, jr;
  unsigned int lmask, rmask;
};

__device__ __forceinline__ RawRow fetch_row(const unsigned int* plane, int x, const LaneCols& lc, int wpr) {
  const unsigned int* row = plane + x * wpr;       // 32-bit index math: (W + 2) * wpr words < 2^31
  RawRow r;
  r.c = __ldcg(row + lc.j);
  r.l = __ldcg(row + lc.jl);
  r.r = __ldcg(row + lc.jr);
  return r;
}

__device__ __forceinline__ RowSums finish_row(const RawRow& raw, const LaneCols& lc) {
  const unsigned int c = raw.c;
  const unsigned int a = __funnelshift_l(raw.l & lc.lmask, c, 1);     // bit i <- cell i-1
  const unsigned int b = __funnelshift_r(c, raw.r & lc.rmask, 1);     // bit i <- cell i+1
  RowSums s;
  s.c = c;
  s.m0 = a ^ b;
  s.m1 = a & b;
  s.s0 = s.m0 ^ c;
  s.s1 = maj3(a, b, c);
  return s;
}

// top(3-sum) + mid(2-sum) + bottom(3-sum) -> 4-bit bit-sliced count (0..8)
__device__ __forceinline__ void add_rows(const RowSums& t, const RowSums& m, const RowSums& b, unsigned int (&z)[4]) {
  z[0] = t.s0 ^ m.m0 ^ b.s0;
  const unsigned int c0 = maj3(t.s0, m.m0, b.s0);
  const unsigned int p = t.s1 ^ m.m1 ^ b.s1;
  const unsigned int q = maj3(t.s1, m.m1, b.s1);
  z[1] = p ^ c0;
  const unsigned int cr = p & c0;
  z[2] = q ^ cr;
  z[3] = q & cr;
}

// (~s & k) | (~(s ^ k) & x): one step of the MSB-first bit-sliced "s < k" chain
__device__ __forceinline__ unsigned int lt_step(unsigned int s, unsigned int k, unsigned int x) {
  return (~s & k) | (~(s ^ k) & x);
}

// same * 840 / o = sum over the set bits b of same of (840 / o) << b.  The (o, b) classes are disjoint
// cell sets, so classes with the same weight share one mask and one popcount: 13 distinct weights.
constexpr int kSegPairs = 13;

// one output row: the 32 cells of a lane's word from the horizontal sums of the rows above (t*), own
// (m*) and below (b*) of both planes.  Returns the unsatisfied-agent mask; adds the cells with an
// occupied neighbour to my_occ and the segregation numerator classes to seg.
__device__ __forceinline__ unsigned int eval_row(const RowSums& to, const RowSums& tt, const RowSums& mo,
                                                 const RowSums& mt, const RowSums& bo, const RowSums& bt,
                                                 const SchellingBitsDev& sb, unsigned int& my_occ,
                                                 unsigned int (&seg)[kSegPairs]) {
  unsigned int o[4], n[4];
  add_rows(to, mo, bo, o);
  add_rows(tt, mt, bt, n);
  const unsigned int A = mo.c, T = mt.c;
  // d = o - n  (n <= o per cell, so no final borrow)
  unsigned int d[4], br;
  d[0] = o[0] ^ n[0];
  br = ~o[0] & n[0];
  d[1] = o[1] ^ n[1] ^ br;
  br = maj3(~o[1], n[1], br);
  d[2] = o[2] ^ n[2] ^ br;
  br = maj3(~o[2], n[2], br);
  d[3] = o[3] ^ n[3] ^ br;
  unsigned int same[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) same[q] = (T & n[q]) | (~T & d[q]);
  // one-hot decode of o = 1..8
  const unsigned int l00 = ~o[1] & ~o[0], l01 = ~o[1] & o[0], l10 = o[1] & ~o[0], l11 = o[1] & o[0];
  const unsigned int h0 = ~o[3] & ~o[2], h1 = ~o[3] & o[2];
  unsigned int is[9];
  is[1] = h0 & l01; is[2] = h0 & l10; is[3] = h0 & l11;
  is[4] = h1 & l00; is[5] = h1 & l01; is[6] = h1 & l10; is[7] = h1 & l11;
  is[8] = o[3];
  // need[o] of every cell as 4 bit-sliced bits, then unsat = agent & (same < need)
  unsigned int K[4] = {0, 0, 0, 0};
#pragma unroll
  for (int k = 1; k <= 8; ++k) {
#pragma unroll
    for (int q = 0; q < 4; ++q) K[q] |= is[k] & sb.need_sel[k][q];
  }
  unsigned int lt = ~same[0] & K[0];
  lt = lt_step(same[1], K[1], lt);
  lt = lt_step(same[2], K[2], lt);
  lt = lt_step(same[3], K[3], lt);
  const unsigned int unsat = A & lt;
  my_occ += __popc(A & (o[0] | o[1] | o[2] | o[3]));
  // segregation numerator: one popcount per distinct weight (840/o) << b
  {
    const unsigned int a1 = A & is[1], a2 = A & is[2], a3 = A & is[3], a4 = A & is[4];
    const unsigned int a5 = A & is[5], a6 = A & is[6], a7 = A & is[7], a8 = A & is[8];
    seg[0] += __popc((a1 & same[0]) | (a2 & same[1]) | (a4 & same[2]) | (a8 & same[3]));   // 840
    seg[1] += __popc((a2 & same[0]) | (a4 & same[1]) | (a8 & same[2]));                     // 420
    seg[2] += __popc((a4 & same[0]) | (a8 & same[1]));                                      // 210
    seg[3] += __popc(a8 & same[0]);                                                         // 105
    seg[4] += __popc((a3 & same[0]) | (a6 & same[1]));                                      // 280
    seg[5] += __popc((a3 & same[1]) | (a6 & same[2]));                                      // 560
    seg[6] += __popc(a6 & same[0]);                                                         // 140
    seg[7] += __popc(a5 & same[0]);                                                         // 168
    seg[8] += __popc(a5 & same[1]);                                                         // 336
    seg[9] += __popc(a5 & same[2]);                                                         // 672
    seg[10] += __popc(a7 & same[0]);                                                        // 120
    seg[11] += __popc(a7 & same[1]);                                                        // 240
    seg[12] += __popc(a7 & same[2]);                                                        // 480
  }
  return unsat;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, JXB_SCH_MINB) schelling_bits_kernel(const SchellingDev sd, const SchellingBitsDev sb,
                                                                  const ModelDev md, int steps) {
  __shared__ unsigned int s_u32[kThreads / 32];
  __shared__ unsigned long long s_u64[kThreads / 32];
  __shared__ unsigned int s_occ[kThreads / 32];
  __shared__ unsigned int s_all[kThreads / 32];
  __shared__ unsigned int s_rk[8];
  __shared__ unsigned int s_prefix, s_total;
  __shared__ unsigned int s_ws[8][kThreads / 32];    // unsatisfied counts per (chunk of a super-chunk, warp)
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kThreads / 32;
  const int B = gridDim.x, b = blockIdx.x;
  Ctrl* ctrl = md.ctrl;
  const TypeDev& t = md.t[0];
  const int W = sd.W, H = sd.H, wpr = sb.wpr, spr = sb.spr, nsub = kWarps / spr;
  // rows owned by this CTA, and by this warp's sub-band inside them
  const int R0 = (int)((long long)W * b / B), R1 = (int)((long long)W * (b + 1) / B);
  const int strip = warp % spr, sub = warp / spr;
  const int rs = R0 + (R1 - R0) * sub / nsub, re = R0 + (R1 - R0) * (sub + 1) / nsub;
  LaneCols lc;
  lc.j = strip * 32 + lane;
  lc.jl = lc.j > 0 ? lc.j - 1 : (sd.periodic ? wpr - 1 : lc.j);
  lc.jr = lc.j + 1 < wpr ? lc.j + 1 : (sd.periodic ? 0 : lc.j);
  lc.lmask = (lc.j > 0 || sd.periodic) ? 0xFFFFFFFFu : 0u;
  lc.rmask = (lc.j + 1 < wpr || sd.periodic) ? 0xFFFFFFFFu : 0u;
  const long long wbeg = (long long)R0 * wpr, wend = (long long)R1 * wpr;       // this CTA's words, in cell order
  const unsigned int e = sd.n_empty;
  const int step0 = ctrl->step_in_run;
  BlkPart* blk_part = (BlkPart*)sd.blk_part;
  const long long words = sd.cells >> 5;

  for (int s = 0; s < steps; ++s) {
    // ------------------------------------------------------------------ phase 1: bit-sliced sweep
    if (tid < 8) {
      const uint32_t* kp = md.keys + (size_t)(step0 + s) * (md.n_types + 1) * 2;
      const Key ck = {kp[0], kp[1]};
      s_rk[tid] = bits_elem<MODE>(ck, tid, 8);
    }
    BlkPart* part = blk_part + (size_t)(s & 1) * B;
    unsigned int my_unsat = 0, my_occ = 0;
    unsigned int seg[kSegPairs];
#pragma unroll
    for (int i = 0; i < kSegPairs; ++i) seg[i] = 0;
    if (rs < re) {
      // rolling window of three rows per plane in R[0..2]: at step i the row above is R[i%3], the own
      // row R[(i+1)%3], and the row below is loaded into R[(i+2)%3]; the loop is unrolled by three
      // so that the rotation is a renaming, not register moves
      RowSums Ro[3], Rt[3];
      Ro[0] = finish_row(fetch_row(sb.occ, rs - 1, lc, wpr), lc);
      Rt[0] = finish_row(fetch_row(sb.t1, rs - 1, lc, wpr), lc);
      Ro[1] = finish_row(fetch_row(sb.occ, rs, lc, wpr), lc);
      Rt[1] = finish_row(fetch_row(sb.t1, rs, lc, wpr), lc);
      RawRow ro = fetch_row(sb.occ, rs + 1, lc, wpr), rt = fetch_row(sb.t1, rs + 1, lc, wpr);
      for (int x0 = rs; x0 < re; x0 += 3) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int x = x0 + i;
          if (x < re) {
            const RowSums& to = Ro[i % 3];
            const RowSums& tt = Rt[i % 3];
            const RowSums& mo = Ro[(i + 1) % 3];
            const RowSums& mt = Rt[(i + 1) % 3];
            Ro[(i + 2) % 3] = finish_row(ro, lc);
            Rt[(i + 2) % 3] = finish_row(rt, lc);
            const RowSums& bo = Ro[(i + 2) % 3];
            const RowSums& bt = Rt[(i + 2) % 3];
            {                      // next row's loads in flight while this row is evaluated (row re+1 of the
              const int xn = min(x + 2, W);   // last band is clamped to the halo)
              ro = fetch_row(sb.occ, xn, lc, wpr);
              rt = fetch_row(sb.t1, xn, lc, wpr);
            }
            const unsigned int unsat = eval_row(to, tt, mo, mt, bo, bt, sb, my_occ, seg);
            sb.umask[x * wpr + lc.j] = unsat;
            my_unsat += __popc(unsat);
          }
        }
      }
    }
    unsigned long long my_num = 0;
    {
      const unsigned int wts[13] = {840, 420, 210, 105, 280, 560, 140, 168, 336, 672, 120, 240, 480};
#pragma unroll
      for (int i = 0; i < kSegPairs; ++i) my_num += (unsigned long long)seg[i] * wts[i];
    }
    {
      const unsigned int a = warp_sum((int)my_unsat), o = warp_sum((int)my_occ);
      unsigned long long n = my_num;
#pragma unroll
      for (int dd = 16; dd > 0; dd >>= 1) n += __shfl_xor_sync(0xffffffffu, n, dd);
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
      for (int dd = 16; dd > 0; dd >>= 1) num += __shfl_xor_sync(0xffffffffu, num, dd);
      __syncthreads();
      if (lane == 0) { s_u32[warp] = before; s_all[warp] = all; s_occ[warp] = occ; s_u64[warp] = num; }
      __syncthreads();
      if (tid == 0) {
        unsigned int p = 0, a = 0, oc = 0;
        unsigned long long nm = 0;
        for (int w = 0; w < kWarps; ++w) { p += s_u32[w]; a += s_all[w]; oc += s_occ[w]; nm += s_u64[w]; }
        s_prefix = p;
        s_total = a;
        if (b == 0) {
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
    // ---------------------------------------------------------- phase 3: moves, in cell order, CTA-local
    // The unsatisfied agents of THIS CTA's rows are U[s_prefix .. s_prefix + my count) in ascending cell order.
    // Every CTA walks its own segment: index j is mover k = piU^-1(j) (if k < m) and goes to slot piE(k) -- the
    // same matching as "mover k is U[piU(k)]", evaluated from the source side, so that the source-side accesses
    // (mask words, the cell payload, the plane words) stream through the CTA's own rows and only the target side
    // (slot, target cell payload, target plane bits) is random.  Targets are cells that were empty when the step
    // began, sources are occupied ones, and every slot belongs to exactly one mover: CTAs need no barrier between
    // the compaction and the moves.  U[j] receives (agent | 1<<31) for a mover and the cell id for an agent that
    // stays -- what the lazy 'satisfied' column needs.
    if (u > 0) {                   // uniform across the grid
      const Feistel fu = make_feistel(u, s_rk), fe = make_feistel(e, s_rk + 4);
      // (a) ordered compaction of the CTA's rows into ITS segment of U (cells ascending), kPre chunks of kThreads
      // mask words at a time: the words of all chunks are fetched together, every warp scans its 32 words of each
      // chunk by shuffles, ONE block barrier publishes the per-(chunk, warp) counts, and every thread then knows
      // where the set bits of its words go.  A super-chunk without an unsatisfied agent costs that one barrier.
      constexpr int kPre = 8;
      unsigned int base = s_prefix;
      for (long long W0 = wbeg; W0 < wend; W0 += (long long)kPre * kThreads) {
        unsigned int pre[kPre], inc[kPre];
        unsigned int any = 0;
#pragma unroll
        for (int c = 0; c < kPre; ++c) {
          const long long w = W0 + (long long)c * kThreads + tid;
          pre[c] = w < wend ? __ldcg(sb.umask + w) : 0u;
        }
#pragma unroll
        for (int c = 0; c < kPre; ++c) {
          const unsigned int cnt = __popc(pre[c]);
          any |= cnt;
          unsigned int v = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t2 = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t2;
          }
          inc[c] = v;
          if (lane == 31) s_ws[c][warp] = v;
        }
        if (__syncthreads_count(any != 0u) != 0) {       // uniform; also orders s_ws writes before the reads below
#pragma unroll
          for (int c = 0; c < kPre; ++c) {
            unsigned int woff = 0, ctot = 0;
#pragma unroll
            for (int ww = 0; ww < kWarps; ++ww) {
              const unsigned int v = s_ws[c][ww];
              if (ww < warp) woff += v;
              ctot += v;
            }
            unsigned int unsat = pre[c];
            unsigned int pu = base + woff + inc[c] - __popc(unsat);
            const unsigned int c0 = (unsigned int)((W0 + (long long)c * kThreads + tid) << 5);
            while (unsat) {
              const int q = __ffs(unsat) - 1;
              unsat &= unsat - 1;
              sd.U[pu++] = c0 + q;
            }
            base += ctot;
          }
        }
        __syncthreads();                                  // s_ws is rewritten by the next super-chunk
      }
      // (b) the moves: the CTA's segment [s_prefix, base) of U, strided over the threads, no barrier inside.  Four
      // entries per thread per iteration: the chain slot / payload -> writes is latency-bound.  (The segment was
      // written by this CTA's own threads: the block barrier above made it visible.)
      constexpr int kMv = JXB_SCH_MV;
      const unsigned int seg_end = base;
      for (unsigned int j0 = s_prefix + tid; j0 < seg_end; j0 += kThreads * kMv) {
        unsigned int src[kMv], jj[kMv], dst[kMv], tw[kMv];
        int2 am[kMv];
        bool mv[kMv];
#pragma unroll
        for (int i = 0; i < kMv; ++i) {
          const unsigned int j = j0 + i * kThreads;
          mv[i] = false; src[i] = 0; jj[i] = 0;
          if (j < seg_end) {
            src[i] = __ldcg(sd.U + j);
            const unsigned int k = feistel_inverse(fu, j);
            mv[i] = k < m;
            if (mv[i]) jj[i] = feistel_permute(fe, k);
          }
        }
#pragma unroll
        for (int i = 0; i < kMv; ++i) {
          dst[i] = 0; tw[i] = 0; am[i] = make_int2(-1, 0);
          if (mv[i]) {
            dst[i] = __ldcg(sd.E + jj[i]);
            am[i] = __ldcg(sb.cell_am + src[i]);
            tw[i] = __ldcg(sb.t1 + (src[i] >> 5));
          }
        }
#pragma unroll
        for (int i = 0; i < kMv; ++i) {
          if (!mv[i]) continue;                           // an agent that stays keeps its cell id in U[j]
          const unsigned int j = j0 + i * kThreads;
          const unsigned int s_ = src[i], d_ = dst[i];
          const unsigned int sbit = 1u << (s_ & 31), dbit = 1u << (d_ & 31);
          const bool ty = (tw[i] & sbit) != 0;
          sd.E[jj[i]] = s_;
          sd.U[j] = (unsigned int)am[i].x | 0x80000000u;
          atomicAnd(sb.occ + (s_ >> 5), ~sbit);
          atomicOr(sb.occ + (d_ >> 5), dbit);
          if (ty) {
            atomicAnd(sb.t1 + (s_ >> 5), ~sbit);
            atomicOr(sb.t1 + (d_ >> 5), dbit);
          }
          sb.cell_am[d_] = make_int2(am[i].x, am[i].y + 1);
          sb.cell_am[s_] = make_int2(-1, 0);
          if (sd.periodic) {         // keep the wrapped halo rows in step
            if (d_ < (unsigned)H) { atomicOr(sb.occ + (d_ >> 5) + words, dbit); if (ty) atomicOr(sb.t1 + (d_ >> 5) + words, dbit); }
            if (d_ >= sd.cells - H) { atomicOr(sb.occ + (long long)(d_ >> 5) - words, dbit); if (ty) atomicOr(sb.t1 + (long long)(d_ >> 5) - words, dbit); }
            if (s_ < (unsigned)H) { atomicAnd(sb.occ + (s_ >> 5) + words, ~sbit); if (ty) atomicAnd(sb.t1 + (s_ >> 5) + words, ~sbit); }
            if (s_ >= sd.cells - H) { atomicAnd(sb.occ + (long long)(s_ >> 5) - words, ~sbit); if (ty) atomicAnd(sb.t1 + (long long)(s_ >> 5) - words, ~sbit); }
          }
        }
      }
    }
    if (m == 0) continue;          // uniform: nobody moved, the planes are unchanged -- no barrier needed
    grid.sync();
  }
}

// cell payload <-> the per-agent API columns (persistent kernel): pack after the grid was (re)built from the
// uploaded positions, unpack when 'position' / 'moves' are read
__global__ void cell_am_pack_kernel(const SchellingDev sd, const SchellingBitsDev sb, const int* moves) {
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < sd.cells; c += (long long)gridDim.x * blockDim.x) {
    const int a = sd.cell_agent[c];
    sb.cell_am[c] = make_int2(a, a >= 0 ? moves[a] : 0);
  }
}

// 'moves' was uploaded between runs: refresh the move counts that travel with the cells (no grid rebuild: the
// empty-cell slot order must survive)
__global__ void cell_am_set_moves_kernel(const SchellingDev sd, const SchellingBitsDev sb, const int* moves) {
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < sd.cells; c += (long long)gridDim.x * blockDim.x) {
    const int a = sb.cell_am[c].x;
    if (a >= 0) sb.cell_am[c].y = moves[a];
  }
}

__global__ void cell_am_unpack_kernel(const SchellingDev sd, const SchellingBitsDev sb, int2* position, int* moves) {
  const int H = sd.H;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < sd.cells; c += (long long)gridDim.x * blockDim.x) {
    const int2 am = __ldcs(sb.cell_am + c);
    if (am.x >= 0) {
      position[am.x] = make_int2((int)(c / H), (int)(c % H));
      moves[am.x] = am.y;
    }
  }
}

// 'satisfied' of the last step from the per-step list the persistent kernel leaves in U: (agent | 1<<31) for an
// agent that moved, the cell of an unsatisfied agent that stayed
__global__ void satisfied_export_packed_kernel(const SchellingDev sd, const SchellingBitsDev sb, const Ctrl* ctrl, unsigned char* sat) {
  const unsigned int u = ctrl->n_unsat;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < u; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int v = sd.U[i];
    const int a = (v >> 31) ? (int)(v & 0x7FFFFFFFu) : sb.cell_am[v].x;
    if (a >= 0) sat[a] = 0;
  }
}

// int8 grid -> bit planes (interior + halo rows), after the grid was (re)built from positions
__global__ void planes_from_ct_kernel(const SchellingDev sd, const SchellingBitsDev sb) {
  const long long words = sd.cells >> 5, wpr = sb.wpr;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x - wpr; w < words + wpr;
       w += (long long)gridDim.x * blockDim.x) {
    long long src = w;                                   // halo rows: wrapped row or empty
    bool empty = false;
    if (w < 0) { if (sd.periodic) src = w + words; else empty = true; }
    if (w >= words) { if (sd.periodic) src = w - words; else empty = true; }
    unsigned int o = 0, t1 = 0;
    if (!empty) {
      const signed char* p = sd.ct + (src << 5);
#pragma unroll 4
      for (int i = 0; i < 32; ++i) {
        const int v = p[i];
        if (v >= 0) { o |= 1u << i; if (v & 1) t1 |= 1u << i; }
      }
    }
    sb.occ[w] = o;
    sb.t1[w] = t1;
  }
}

// bit planes -> int8 grid interior (env['grid'] export after a run)
__global__ void ct_from_planes_kernel(const SchellingDev sd, const SchellingBitsDev sb) {
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < sd.cells; c += (long long)gridDim.x * blockDim.x) {
    const unsigned int bit = 1u << (c & 31);
    const bool o = sb.occ[c >> 5] & bit, ty = sb.t1[c >> 5] & bit;
    sd.ct[c] = o ? (signed char)(ty ? 1 : 0) : (signed char)-1;
  }
}

}  // namespace jxb
