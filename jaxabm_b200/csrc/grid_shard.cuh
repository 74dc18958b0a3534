// grid_shard.cuh -- ONE Schelling grid decomposed into row bands over the GPUs of a box (SURVEY.md
// 8(e) "Grid: row blocks + halo"; one process per GPU, peers' receive areas mapped over NVLink by
// CUDA IPC).  Results are bit-identical to the single-GPU kernels (schelling_bits.cuh).
//
// Everything a step touches is PARTITIONED -- a rank only walks the movers of its own band, the slots of its own
// range and the records addressed to its own rows; every remote access is a posted store:
//   cells   rank r owns rows [X0, X1): the bit planes occ / t1 (plus one halo row each side, kept
//           coherent by the movers' records) and the cell payload cell_am (agent, moves);
//   U       the ordered list of unsatisfied cells: the bands' lists in rank order, so rank r holds
//           U[prefix[r] .. prefix[r+1]) -- its own rows, compacted CTA-locally;
//   E       the empty-cell slots: rank q holds slots [q * eper, (q+1) * eper) inside its receive area.
//
// A step is five launches per rank, graph-captured, with no host round trip and no NCCL call:
//   1 sweep    bit-sliced neighbour counts of the band (eval_row of schelling_bits.cuh): unsatisfied
//              mask + exact integer partials per CTA.
//   2 counts   one CTA: folds the partials, stores the band's counts + flag A into every rank's
//              header, waits for the world's flags, folds the counts in rank order (exact integers
//              -> the metrics row is identical everywhere) and leaves u, m = min(u, e) and the
//              ranks' prefixes in U.
//   3 moveout  every CTA compacts the unsatisfied cells of ITS rows into its segment of U and walks
//              them in cell order: entry j is mover k = piU^-1(j) (if k < m) and takes slot piE(k) --
//              the single-GPU matching, evaluated from the source side.  The source cell is cleared
//              locally (and, for a boundary row, in the neighbours' halo copies by a record); the
//              mover leaves as a 16-byte request (slot, source cell, agent | type<<31, moves) stored
//              into segment [parity][me] of the SLOT OWNER's receive area; the last CTA releases the
//              per-target counts + flag B.
//   4 forward  (every CTA first waits for the world's B flags in its OWN area.)  The slot owner reads
//              each request's slot (the target cell), rewrites it with the source cell, and forwards
//              (target cell, agent | type<<31, moves + 1, set) to the rank that owns the target row --
//              plus a plane-only copy to the ranks that hold that row as a halo; the last CTA
//              releases flag C.
//   5 apply    (waits for the C flags.)  Every received record sets / clears the plane bits wherever
//              this rank keeps that row (band, halo, wrapped halo of a periodic grid) and, for a cell
//              of the band, writes the payload.
// (A first version let the mover's rank read and rewrite the slot in place over NVLink: half of the
// slot reads at 2 GPUs being remote loads cost more than the halved work saved -- measured,
// profiles/r02_grid_bands_v2.txt -- so the slot owner became a party of its own.)
// Segments and counts are double-buffered by step parity: a rank runs at most one exchange ahead of
// a peer that is still reading the previous one.
#pragma once
#include "common.cuh"
#include "schelling.cuh"
#include "schelling_bits.cuh"

namespace jxb {

struct GridXchgHdr {
  unsigned int flagA[2][kMaxPeers];       // step tag: rank p's band counts of this parity are in cntA
  unsigned int cntA[2][kMaxPeers][4];     // rank p's band: #unsatisfied, #with a neighbour, numerator lo / hi
  unsigned int flagB[2][kMaxPeers];       // step tag: rank p's requests of this parity are complete
  unsigned int cntB[2][kMaxPeers];        // how many requests rank p sent here
  unsigned int flagC[2][kMaxPeers];       // step tag: rank p's records of this parity are complete
  unsigned int cntC[2][kMaxPeers][2];     // [1]: how many halo records rank p left at the back of its segment ([0] is not used:
                                          // the cell records are counted per chunk, gs_chunk_cnt)
  unsigned int err;                       // 1: a peer's flag did not arrive within the spin budget, 2: a segment overflowed
  unsigned int pad[256 - 20 * kMaxPeers - 1];
};
static_assert(sizeof(GridXchgHdr) == 1024, "receive-area header is 1 KiB");

struct GridStepInfo {
  unsigned int tag, par, u, m;
  int key_row;                            // row of the run's key table that belongs to this step
  unsigned int ticket[2];                 // last-CTA election of the moveout / forward kernel
  unsigned int prefix[kMaxPeers + 1];     // rank q's unsatisfied cells are U[prefix[q] .. prefix[q+1])
};

struct GridShardDev {
  int rank, world;
  int X0, X1;                             // rows owned by this rank
  int xb[kMaxPeers + 1];                  // rank q owns rows [xb[q], xb[q+1])
  unsigned int cap, halo_cap;             // records per (parity, rank) segment; its last halo_cap entries take the halo records
  unsigned int eper;                      // empty-cell slots per rank = requests per (parity, rank) segment
  unsigned int nfwd, fchunk;              // CTAs of the forward kernel (the same on every rank); each owns a chunk of fchunk
                                          // records in its segment of every target: cap = nfwd * fchunk + halo_cap
  unsigned long long slot_off, req_off, rec_off;    // byte offsets in a receive area: [header][chunk counts u32[2][world][nfwd]]
                                          // [slots][request segments][record segments]
  unsigned char* peer[kMaxPeers];         // every rank's receive area as mapped here (own = local)
  unsigned char* self;                    // == peer[rank]
  GridStepInfo* info;
  BlkPart* part;                          // [blocks] per-CTA partials of the sweep
  unsigned int* sendcnt;                  // [3][kMaxPeers] appended for each target this step: row 1 halo records, row 2 requests
                                          // (row 0 is not used)
  int blocks;                             // CTAs of the sweep AND the moveout kernel (same row split)
};

__device__ __forceinline__ GridXchgHdr* gs_hdr(unsigned char* base) { return (GridXchgHdr*)base; }
// peer[p] for a run-time p without spilling the kernel-parameter array to local memory
__device__ __forceinline__ unsigned char* gs_peer(const GridShardDev& gs, int p) {
  unsigned char* r = gs.peer[0];
#pragma unroll
  for (int i = 1; i < kMaxPeers; ++i)
    if (p == i) r = gs.peer[i];
  return r;
}
__device__ __forceinline__ int gs_owner(const GridShardDev& gs, int x) {
  int q = 0;
#pragma unroll
  for (int i = 1; i < kMaxPeers; ++i)
    if (i < gs.world && x >= gs.xb[i]) q = i;
  return q;
}
__device__ __forceinline__ uint4* gs_seg(unsigned char* base, const GridShardDev& gs, unsigned int par, int from) {
  return (uint4*)(base + gs.rec_off) + ((size_t)par * gs.world + from) * gs.cap;
}
// how many records CTA c of rank `from`'s forward kernel left in its chunk of my segment
__device__ __forceinline__ unsigned int* gs_chunk_cnt(unsigned char* base, const GridShardDev& gs, unsigned int par, int from) {
  return (unsigned int*)(base + sizeof(GridXchgHdr)) + ((size_t)par * gs.world + from) * gs.nfwd;
}
__device__ __forceinline__ unsigned int* gs_slots(unsigned char* base, const GridShardDev& gs) {
  return (unsigned int*)(base + gs.slot_off);
}
__device__ __forceinline__ uint4* gs_req(unsigned char* base, const GridShardDev& gs, unsigned int par, int from) {
  return (uint4*)(base + gs.req_off) + ((size_t)par * gs.world + from) * gs.eper;
}

// ------------------------------------------------------------------------------------ 1 sweep
__global__ void __launch_bounds__(kThreads, 2) grid_shard_sweep_kernel(const SchellingDev sd, const SchellingBitsDev sb,
                                                                    const GridShardDev gs) {
  constexpr int kWarps = kThreads / 32;
  __shared__ unsigned int s_u32[kWarps], s_occ[kWarps];
  __shared__ unsigned long long s_u64[kWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = gridDim.x, b = blockIdx.x;
  const int W = sd.W, wpr = sb.wpr;
  const int nrows = gs.X1 - gs.X0;
  const int R0 = gs.X0 + (int)((long long)nrows * b / B), R1 = gs.X0 + (int)((long long)nrows * (b + 1) / B);
  const int strips = (wpr + 31) >> 5;
  const int nsub = strips >= kWarps ? 1 : kWarps / strips;
  unsigned int my_unsat = 0, my_occ = 0;
  unsigned int seg[kSegPairs];
#pragma unroll
  for (int i = 0; i < kSegPairs; ++i) seg[i] = 0;
  for (int task = warp; task < strips * nsub; task += kWarps) {
    const int strip = task % strips, sub = task / strips;
    const int rs = R0 + (R1 - R0) * sub / nsub, re = R0 + (R1 - R0) * (sub + 1) / nsub;
    LaneCols lc;
    lc.j = strip * 32 + lane;
    if (lc.j >= wpr || rs >= re) continue;        // per-lane: the row loop holds no warp collectives
    lc.jl = lc.j > 0 ? lc.j - 1 : (sd.periodic ? wpr - 1 : lc.j);
    lc.jr = lc.j + 1 < wpr ? lc.j + 1 : (sd.periodic ? 0 : lc.j);
    lc.lmask = (lc.j > 0 || sd.periodic) ? 0xFFFFFFFFu : 0u;
    lc.rmask = (lc.j + 1 < wpr || sd.periodic) ? 0xFFFFFFFFu : 0u;
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
          {
            const int xn = min(x + 2, W);          // prefetch; rows beyond the halo are loaded but never used
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
  __syncwarp();
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
    gs.part[b] = p;
  }
}

// ------------------------------------------------------------------------------------ 2 counts
__global__ void __launch_bounds__(kThreads) grid_shard_counts_kernel(const SchellingDev sd, const GridShardDev gs, const ModelDev md) {
  constexpr int kWarps = kThreads / 32;
  __shared__ unsigned int s_u32[kWarps], s_occ[kWarps];
  __shared__ unsigned long long s_u64[kWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  Ctrl* ctrl = md.ctrl;
  const TypeDev& t = md.t[0];
  const unsigned int tag = (unsigned int)(ctrl->time_step + 1);
  const unsigned int par = tag & 1u;
  {
    unsigned int all = 0, occ = 0;
    unsigned long long num = 0;
    for (int i = tid; i < gs.blocks; i += kThreads) {
      const uint4 raw = __ldcg((const uint4*)(gs.part + i));
      all += raw.x;
      occ += raw.y;
      num += ((unsigned long long)raw.w << 32) | raw.z;
    }
    all = warp_sum((int)all);
    occ = warp_sum((int)occ);
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) num += __shfl_xor_sync(0xffffffffu, num, dd);
    if (lane == 0) { s_u32[warp] = all; s_occ[warp] = occ; s_u64[warp] = num; }
    __syncthreads();
  }
  if (warp != 0) return;
  GridXchgHdr* h = gs_hdr(gs.self);
  unsigned int c0 = 0, c1 = 0;
  unsigned long long num = 0;
  if (lane < gs.world) {
    unsigned int a = 0, oc = 0;
    unsigned long long nm = 0;
    for (int w = 0; w < kWarps; ++w) { a += s_u32[w]; oc += s_occ[w]; nm += s_u64[w]; }
    // my band's counts into rank `lane`'s header, then the flag (release: the counts precede it)
    GridXchgHdr* ph = gs_hdr(gs_peer(gs, lane));
    unsigned int* c = ph->cntA[par][gs.rank];
    c[0] = a;
    c[1] = oc;
    c[2] = (unsigned int)nm;
    c[3] = (unsigned int)(nm >> 32);
    st_release_sys(&ph->flagA[par][gs.rank], tag);
    if (!*(volatile unsigned int*)&h->err) {
      const long long t0 = clock64();
      while (ld_acquire_sys(&h->flagA[par][lane]) != tag) {
        if (clock64() - t0 > (20ll << 30)) { h->err = 1u; break; }     // ~10 s: a peer is gone
      }
    }
    c0 = __ldcv(&h->cntA[par][lane][0]);
    c1 = __ldcv(&h->cntA[par][lane][1]);
    num = ((unsigned long long)__ldcv(&h->cntA[par][lane][3]) << 32) | __ldcv(&h->cntA[par][lane][2]);
  }
  __syncwarp();
  unsigned int inc = c0;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  const unsigned int u = __shfl_sync(0xffffffffu, inc, 31);
  const unsigned int oc = (unsigned int)warp_sum((int)c1);
#pragma unroll
  for (int dd = 16; dd > 0; dd >>= 1) num += __shfl_xor_sync(0xffffffffu, num, dd);
  GridStepInfo* info = gs.info;
  if (lane <= gs.world && lane <= kMaxPeers) info->prefix[lane] = lane < gs.world ? inc - c0 : u;
  if (lane == 0) {
    const unsigned int e = sd.n_empty;
    const unsigned int m = u < e ? u : e;
    const long long ts = ctrl->time_step + 1;
    info->tag = tag; info->par = par; info->u = u; info->m = m;
    info->key_row = ctrl->step_in_run;
    ctrl->total_moves += m;
    ctrl->n_unsat = u;
    ctrl->n_moved = m;
    const double segv = (double)(float)((double)num / 840.0 / (double)(oc ? oc : 1));
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

// ------------------------------------------------------------------------------------ 3 moveout
__device__ __forceinline__ void gs_plane_write(const SchellingBitsDev& sb, long long w, unsigned int bit, bool ty, bool set) {
  if (set) {
    atomicOr(sb.occ + w, bit);
    if (ty) atomicOr(sb.t1 + w, bit);
  } else {
    atomicAnd(sb.occ + w, ~bit);
    if (ty) atomicAnd(sb.t1 + w, ~bit);
  }
}

// flip cell c (row x) in every plane row this rank keeps coherent: band + halo rows, and the wrapped
// halo rows of a periodic grid when the band touches the grid edge
__device__ __forceinline__ void gs_flip(const SchellingDev& sd, const SchellingBitsDev& sb, const GridShardDev& gs,
                                        unsigned int c, int x, bool ty, bool set) {
  const unsigned int bit = 1u << (c & 31);
  const long long w = c >> 5, words = sd.cells >> 5;
  if (x >= gs.X0 - 1 && x <= gs.X1) gs_plane_write(sb, w, bit, ty, set);
  if (sd.periodic) {
    if (x == 0 && gs.X1 == sd.W) gs_plane_write(sb, w + words, bit, ty, set);
    if (x == sd.W - 1 && gs.X0 == 0) gs_plane_write(sb, w - words, bit, ty, set);
  }
}

// owner of row x and its band [lo, hi)
__device__ __forceinline__ int gs_owner3(const GridShardDev& gs, int x, int& lo, int& hi) {
  int q = 0;
  lo = gs.xb[0]; hi = gs.xb[1];
#pragma unroll
  for (int i = 1; i < kMaxPeers; ++i)
    if (i < gs.world && x >= gs.xb[i]) { q = i; lo = gs.xb[i]; hi = gs.xb[i + 1]; }
  return q;
}

// the ranks other than `own` that keep row x as a halo row: the owners of the rows above and below it
__device__ __forceinline__ void gs_halo_ranks(const SchellingDev& sd, const GridShardDev& gs, int x, int own, int& ha, int& hb) {
  int xa = x - 1, xc = x + 1;
  if (sd.periodic) { if (xa < 0) xa = sd.W - 1; if (xc >= sd.W) xc = 0; }
  ha = xa >= 0 ? gs_owner(gs, xa) : own;
  hb = xc < sd.W ? gs_owner(gs, xc) : own;
  if (ha == own) ha = -1;
  if (hb == own || hb == ha) hb = -1;
}

// a record that only concerns a halo copy: rare (boundary rows), appended from the back of the segment
__device__ __forceinline__ void gs_send_halo(const GridShardDev& gs, unsigned int par, int p, uint4 rec) {
  const unsigned int i = atomicAdd(gs.sendcnt + kMaxPeers + p, 1u);
  if (i < gs.halo_cap) gs_seg(gs_peer(gs, p), gs, par, gs.rank)[gs.cap - 1u - i] = rec;
  else gs_hdr(gs.self)->err = 2u;
}

#ifndef JXB_GS_MINB
#define JXB_GS_MINB 2     // resident CTAs per SM the moveout kernel's register allocation aims at
#endif

template <int MODE>
__global__ void __launch_bounds__(kThreads, JXB_GS_MINB) grid_shard_moveout_kernel(const SchellingDev sd, const SchellingBitsDev sb,
                                                                      const GridShardDev gs, const ModelDev md) {
  constexpr int kWarps = kThreads / 32;
  __shared__ unsigned int s_u32[kWarps];
  __shared__ unsigned int s_ws[8][kWarps];
  __shared__ unsigned int s_rk[8];
  __shared__ unsigned int s_cnt[2][kMaxPeers], s_base[2][kMaxPeers];
  __shared__ uint4* s_seg[kMaxPeers];              // my request segment of this parity in rank q's receive area
  __shared__ unsigned int s_prefix, s_last;
  const GridStepInfo* info = gs.info;
  const unsigned int u = info->u, m = info->m, par = info->par;
  if (u == 0) return;                    // uniform over the world: nobody sends, nobody waits
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = gridDim.x, b = blockIdx.x;
  const int wpr = sb.wpr, H = sd.H;
  const int nrows = gs.X1 - gs.X0;
  const int R0 = gs.X0 + (int)((long long)nrows * b / B), R1 = gs.X0 + (int)((long long)nrows * (b + 1) / B);
  const long long wbeg = (long long)R0 * wpr, wend = (long long)R1 * wpr;
  if (tid < 8) {
    const uint32_t* kp = md.keys + (size_t)info->key_row * (md.n_types + 1) * 2;
    const Key ck = {kp[0], kp[1]};
    s_rk[tid] = bits_elem<MODE>(ck, tid, 8);
  }
  if (tid < kMaxPeers) {
    s_seg[tid] = gs_req(gs_peer(gs, tid < gs.world ? tid : 0), gs, par, gs.rank);
    s_cnt[0][tid] = 0u; s_cnt[1][tid] = 0u;
  }
  {
    unsigned int before = 0;
    for (int i = tid; i < b; i += kThreads) before += __ldcg(&gs.part[i].unsat);
    before = warp_sum((int)before);
    if (lane == 0) s_u32[warp] = before;
    __syncthreads();
    if (tid == 0) {
      unsigned int p = info->prefix[gs.rank];
      for (int w = 0; w < kWarps; ++w) p += s_u32[w];
      s_prefix = p;
    }
    __syncthreads();
  }
  // (a) ordered compaction of the CTA's rows into ITS segment of U (as in schelling_bits_kernel)
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
    if (__syncthreads_count(any != 0u) != 0) {
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
    __syncthreads();
  }
  // (b) the movers of the CTA's segment [s_prefix, base), in cell order: the source side of the move happens here,
  // the rest travels as a request to the owner of the slot.  U[j] receives (agent | 1<<31) for a mover and keeps the
  // cell id of an agent that stays -- what the lazy 'satisfied' column needs.
  if (m > 0) {
    const Feistel fu = make_feistel(u, s_rk), fe = make_feistel(sd.n_empty, s_rk + 4);
    constexpr int kMv = 4;
    const unsigned int seg_end = base;
    const int me = gs.rank;
    const int hs = (H & (H - 1)) == 0 ? __ffs(H) - 1 : -1;       // rows by shift when the row length is a power of two
    const long long words = sd.cells >> 5;
    int it = 0;
    for (unsigned int b0 = s_prefix; b0 < seg_end; b0 += kThreads * kMv, it ^= 1) {     // uniform trip count: barriers inside
      unsigned int src[kMv], k[kMv], tw[kMv], li[kMv];
      int2 am[kMv];
      const unsigned int j0 = b0 + tid;
      const int cnt = j0 < seg_end ? (int)min((unsigned int)kMv, (seg_end - j0 + kThreads - 1) / kThreads) : 0;
#pragma unroll
      for (int i = 0; i < kMv; ++i) src[i] = i < cnt ? __ldcg(sd.U + j0 + i * kThreads) : 0u;
      feistel_inverse_strided<kMv>(fu, j0, kThreads, cnt, k);
      unsigned int mv = 0;               // bit i: entry i moves
#pragma unroll
      for (int i = 0; i < kMv; ++i)
        if (i < cnt && k[i] < m) mv |= 1u << i;
#pragma unroll
      for (int i = 0; i < kMv; ++i) {
        tw[i] = 0; am[i] = make_int2(-1, 0);
        if ((mv >> i) & 1u) {
          am[i] = __ldcg(sb.cell_am + src[i]);
          tw[i] = __ldcg(sb.t1 + (src[i] >> 5));
        }
      }
      feistel_permute_masked<kMv>(fe, mv, k);         // k[i] is now the slot index of mover i
#pragma unroll
      for (int i = 0; i < kMv; ++i) {
        li[i] = 0;
        if (!((mv >> i) & 1u)) continue;
        const unsigned int s_ = src[i];
        const unsigned int sbit = 1u << (s_ & 31);
        const bool ty = (tw[i] & sbit) != 0;
        const int xs = hs >= 0 ? (int)(s_ >> hs) : (int)(s_ / (unsigned int)H);
        sd.U[j0 + i * kThreads] = (unsigned int)am[i].x | 0x80000000u;
        atomicAnd(sb.occ + (s_ >> 5), ~sbit);            // the source is a cell of my band
        if (ty) atomicAnd(sb.t1 + (s_ >> 5), ~sbit);
        sb.cell_am[s_] = make_int2(-1, 0);
        if (xs == gs.X0 || xs == gs.X1 - 1) {            // boundary row: my wrapped halo copy, the neighbours' halo copies
          if (sd.periodic) {
            if (xs == 0 && gs.X1 == sd.W) gs_plane_write(sb, (long long)(s_ >> 5) + words, sbit, ty, false);
            if (xs == sd.W - 1 && gs.X0 == 0) gs_plane_write(sb, (long long)(s_ >> 5) - words, sbit, ty, false);
          }
          int ha, hb;
          gs_halo_ranks(sd, gs, xs, me, ha, hb);
          const uint4 rec = make_uint4(s_, ty ? 0x80000000u : 0u, 0u, 0u);
          if (ha >= 0) gs_send_halo(gs, par, ha, rec);
          if (hb >= 0) gs_send_halo(gs, par, hb, rec);
        }
        li[i] = atomicAdd(&s_cnt[it][k[i] / gs.eper], 1u);
      }
      __syncthreads();
      if (tid < kMaxPeers) {
        const unsigned int c = s_cnt[it][tid];
        s_base[it][tid] = c ? atomicAdd(gs.sendcnt + 2 * kMaxPeers + tid, c) : 0u;
        s_cnt[it ^ 1][tid] = 0u;          // the other buffer: read two barriers ago, next written after the barrier below
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kMv; ++i) {
        if (!((mv >> i) & 1u)) continue;
        const unsigned int q = k[i] / gs.eper;
        const unsigned int idx = s_base[it][q] + li[i];
        const uint4 req = make_uint4(k[i] - q * gs.eper, src[i], (unsigned int)am[i].x | (((tw[i] >> (src[i] & 31)) & 1u) << 31),
                                     (unsigned int)am[i].y);
        if (idx < gs.eper) s_seg[q][idx] = req;
        else gs_hdr(gs.self)->err = 2u;
      }
    }
    // the last CTA to finish publishes the counts + flag: the requests of ALL CTAs precede it (the block barrier
    // orders every thread's stores before thread 0's system-scope fence, which is cumulative)
    __syncthreads();
    if (tid == 0) {
      __threadfence_system();
      s_last = (atomicAdd(&gs.info->ticket[0], 1u) == (unsigned int)(B - 1)) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
      if (tid == 0) gs.info->ticket[0] = 0u;
      __threadfence_system();
      if (tid < gs.world) {
        GridXchgHdr* ph = gs_hdr(gs_peer(gs, tid));
        ph->cntB[par][me] = atomicExch(gs.sendcnt + 2 * kMaxPeers + tid, 0u);
        st_release_sys(&ph->flagB[par][me], info->tag);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ waiting for the peers
// Every CTA of the consuming kernel waits for the world's flags itself (thread p spins on rank p's flag in this
// rank's OWN area, ld.acquire.sys) and leaves the prefix of the senders' counts in s_rcv[0 .. world]: no separate
// wait launch.  CH = 0: the requests (flag B), 1: the records (flag C; the cell records are counted per chunk by
// the senders, the halo records here).
template <int CH>
__device__ __forceinline__ void gs_wait(const GridShardDev& gs, const GridStepInfo* info, unsigned int* s_rcv) {
  const int tid = threadIdx.x;
  const unsigned int tag = info->tag, par = info->par;
  GridXchgHdr* h = gs_hdr(gs.self);
  if (tid < gs.world) {
    const unsigned int* flag = CH ? &h->flagC[par][tid] : &h->flagB[par][tid];
    if (!*(volatile unsigned int*)&h->err) {
      const long long t0 = clock64();
      while (ld_acquire_sys(flag) != tag) {
        if (clock64() - t0 > (20ll << 30)) { h->err = 1u; break; }     // ~10 s: a peer is gone
      }
    }
    s_rcv[tid + 1] = CH ? __ldcv(&h->cntC[par][tid][1]) : __ldcv(&h->cntB[par][tid]);
  }
  __syncthreads();
  if (tid == 0) {
    unsigned int acc = 0;
    s_rcv[0] = 0u;
    for (int q = 1; q <= gs.world; ++q) { acc += s_rcv[q]; s_rcv[q] = acc; }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------ 4 forward
// the slot owner's part: slot -> target cell, slot <- source cell, the mover goes on to the owner of the target row.
// CTA c serves a contiguous block of the received requests and appends its records to ITS chunk of my segment in
// every target's receive area: positions come from shared-memory counters (one atomic per warp and target), so the
// loop holds no barrier and no global atomic -- it is a chain of two dependent random accesses per request and
// needs every warp running free to hide them.
__global__ void __launch_bounds__(kThreads, 4) grid_shard_forward_kernel(const SchellingDev sd, const GridShardDev gs) {
  __shared__ unsigned int s_rcv[kMaxPeers + 1];
  __shared__ unsigned int s_pos[kMaxPeers];
  __shared__ uint4* s_chunk[kMaxPeers];            // my chunk of my record segment of this parity in rank p's receive area
  __shared__ unsigned int s_last;
  const GridStepInfo* info = gs.info;
  if (info->m == 0) return;              // uniform over the world
  const int tid = threadIdx.x, lane = tid & 31;
  const unsigned int par = info->par;
  const int me = gs.rank, H = sd.H, c = blockIdx.x;
  if (tid < kMaxPeers) {
    s_chunk[tid] = gs_seg(gs_peer(gs, tid < gs.world ? tid : 0), gs, par, me) + (size_t)c * gs.fchunk;
    s_pos[tid] = 0u;
  }
  gs_wait<0>(gs, info, s_rcv);
  const unsigned int total = s_rcv[gs.world];
  const unsigned int per = (total + gridDim.x - 1) / gridDim.x;
  const unsigned int lo = min(total, (unsigned int)c * per), hi = min(total, lo + per);
  unsigned int* slots = gs_slots(gs.self, gs);
  const int hs = (H & (H - 1)) == 0 ? __ffs(H) - 1 : -1;
  constexpr int kRq = 4;
  for (unsigned int c0 = lo; c0 < hi; c0 += kThreads * kRq) {      // uniform per CTA: the warps stay converged for the votes
    uint4 rq[kRq];
    unsigned int dst[kRq];
    bool ok[kRq];
#pragma unroll
    for (int r = 0; r < kRq; ++r) {
      const unsigned int i = c0 + r * kThreads + tid;
      ok[r] = i < hi;
      rq[r] = make_uint4(0u, 0u, 0u, 0u);
      if (ok[r]) {
        int sq = 0;
        while (sq + 1 < gs.world && i >= s_rcv[sq + 1]) ++sq;
        rq[r] = __ldcs(gs_req(gs.self, gs, par, sq) + (i - s_rcv[sq]));
      }
    }
#pragma unroll
    for (int r = 0; r < kRq; ++r) dst[r] = ok[r] ? __ldcg(slots + rq[r].x) : 0u;
#pragma unroll
    for (int r = 0; r < kRq; ++r) {
      int p = -1;
      const unsigned int d_ = dst[r];
      if (ok[r]) {
        slots[rq[r].x] = rq[r].y;
        const int xd = hs >= 0 ? (int)(d_ >> hs) : (int)(d_ / (unsigned int)H);
        int blo, bhi;
        p = gs_owner3(gs, xd, blo, bhi);
        if (xd == blo || xd == bhi - 1) {                // the target row is somebody's halo row
          int ha, hb;
          gs_halo_ranks(sd, gs, xd, p, ha, hb);
          const uint4 rec = make_uint4(d_, rq[r].z, rq[r].w + 1u, 1u);
          if (ha >= 0) gs_send_halo(gs, par, ha, rec);
          if (hb >= 0) gs_send_halo(gs, par, hb, rec);
        }
      }
      // one shared-memory atomic per (warp, target): the lanes of a group take consecutive positions
      const unsigned int grp = __match_any_sync(0xffffffffu, p);
      const int leader = __ffs(grp) - 1;
      unsigned int pos = 0;
      if (lane == leader && p >= 0) pos = atomicAdd(&s_pos[p], (unsigned int)__popc(grp));
      pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(grp & ((1u << lane) - 1u));
      if (p >= 0) {
        if (pos < gs.fchunk) s_chunk[p][pos] = make_uint4(d_, rq[r].z, rq[r].w + 1u, 1u);
        else gs_hdr(gs.self)->err = 2u;
      }
    }
  }
  __syncthreads();
  if (tid < gs.world) gs_chunk_cnt(gs_peer(gs, tid), gs, par, me)[c] = s_pos[tid];
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();
    s_last = (atomicAdd(&gs.info->ticket[1], 1u) == gridDim.x - 1u) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    if (tid == 0) gs.info->ticket[1] = 0u;
    __threadfence_system();
    if (tid < gs.world) {
      GridXchgHdr* ph = gs_hdr(gs_peer(gs, tid));
      ph->cntC[par][me][1] = atomicExch(gs.sendcnt + kMaxPeers + tid, 0u);
      st_release_sys(&ph->flagC[par][me], info->tag);
    }
  }
}

// ------------------------------------------------------------------------------------ 5 apply
__device__ __forceinline__ void gs_apply_record(const SchellingDev& sd, const SchellingBitsDev& sb, const GridShardDev& gs,
                                                const uint4 rec, int hs) {
  const unsigned int c = rec.x;
  const bool ty = (rec.y >> 31) != 0, set = rec.w != 0u;
  const int x = hs >= 0 ? (int)(c >> hs) : (int)(c / (unsigned int)sd.H);
  gs_flip(sd, sb, gs, c, x, ty, set);
  if (set && x >= gs.X0 && x < gs.X1) sb.cell_am[c] = make_int2((int)(rec.y & 0x7FFFFFFFu), (int)rec.z);
}

__global__ void __launch_bounds__(256) grid_shard_apply_kernel(const SchellingDev sd, const SchellingBitsDev sb, const GridShardDev gs) {
  __shared__ unsigned int s_rcv[kMaxPeers + 1];
  const GridStepInfo* info = gs.info;
  if (info->m == 0) return;
  const int tid = threadIdx.x;
  gs_wait<1>(gs, info, s_rcv);
  const unsigned int par = info->par;
  const int H = sd.H;
  const int hs = (H & (H - 1)) == 0 ? __ffs(H) - 1 : -1;
  constexpr int kRec = 4;                // records in flight per thread: the loop is a chain of dependent random accesses
  // the cell records: one (sender, chunk) pair per CTA at a time
  const int pairs = gs.world * (int)gs.nfwd;
  for (int pi = blockIdx.x; pi < pairs; pi += gridDim.x) {
    const int from = pi / (int)gs.nfwd, c = pi - from * (int)gs.nfwd;
    const unsigned int n = __ldcv(gs_chunk_cnt(gs.self, gs, par, from) + c);
    const uint4* chunk = gs_seg(gs.self, gs, par, from) + (size_t)c * gs.fchunk;
    for (unsigned int i0 = tid; i0 < n; i0 += 256 * kRec) {
      uint4 rec[kRec];
#pragma unroll
      for (int r = 0; r < kRec; ++r) {
        const unsigned int i = i0 + r * 256;
        rec[r] = i < n ? __ldcs(chunk + i) : make_uint4(0u, 0u, 0u, 2u);
      }
#pragma unroll
      for (int r = 0; r < kRec; ++r)
        if (rec[r].w <= 1u) gs_apply_record(sd, sb, gs, rec[r], hs);
    }
  }
  // the halo records (boundary rows only)
  const unsigned int total = s_rcv[gs.world];
  for (unsigned int i = blockIdx.x * blockDim.x + tid; i < total; i += gridDim.x * blockDim.x) {
    int sq = 0;
    while (sq + 1 < gs.world && i >= s_rcv[sq + 1]) ++sq;
    gs_apply_record(sd, sb, gs, __ldcg(gs_seg(gs.self, gs, par, sq) + (gs.cap - 1u - (i - s_rcv[sq]))), hs);
  }
}

// ------------------------------------------------------------------------------------ setup / export
// 'position' / 'moves' of the agents sitting in the band from the cell payload; everybody else keeps the
// fill (-1 / 0: the host combines the ranks with max / sum)
__global__ void grid_shard_unpack_kernel(const SchellingDev sd, const SchellingBitsDev sb, const GridShardDev gs, int2* pos, int* moves) {
  const long long c_begin = (long long)gs.X0 * sd.H, c_end = (long long)gs.X1 * sd.H;
  for (long long c = c_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; c < c_end;
       c += (long long)gridDim.x * blockDim.x) {
    const int2 am = __ldcs(sb.cell_am + c);
    if (am.x >= 0) {
      pos[am.x] = make_int2((int)(c / sd.H), (int)(c % sd.H));
      moves[am.x] = am.y;
    }
  }
}

// 'satisfied' of the last step: the band's unsatisfied agents are this rank's segment of U -- (agent | 1<<31)
// for one that moved, the cell of one that stayed; everybody else keeps the 1 fill (the host combines with min)
__global__ void grid_shard_export_satisfied_kernel(const SchellingBitsDev sb, const SchellingDev sd, const GridShardDev gs,
                                                   unsigned char* sat) {
  const GridStepInfo* info = gs.info;
  const unsigned int lo = info->prefix[gs.rank], hi = info->prefix[gs.rank + 1];
  for (unsigned int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
    const unsigned int v = sd.U[i];
    const int a = (v >> 31) ? (int)(v & 0x7FFFFFFFu) : sb.cell_am[v].x;
    if (a >= 0) sat[a] = 0;
  }
}

}  // namespace jxb
