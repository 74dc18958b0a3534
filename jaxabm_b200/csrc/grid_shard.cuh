// grid_shard.cuh -- ONE Schelling grid decomposed into row bands over the GPUs of a box (SURVEY.md
// 8(e) "Grid: row blocks + halo"; one process per GPU, peers' receive areas mapped over NVLink by
// CUDA IPC).  Results are bit-identical to the single-GPU kernels (schelling_bits.cuh).
//
// Rank r owns rows [X0, X1) of the W x H grid.  Every rank keeps the global-indexed arrays of the
// single-GPU engine (bit planes occ / t1, cell_agent, the replicated empty-cell slots E), but only
// rows [X0-1, X1] of the planes (the band + one halo row each side) and rows [X0, X1) of cell_agent
// are kept coherent and touched per step.
//
// A step is four launches per rank, graph-captured, with no host round trip and no NCCL call:
//   1 sweep    bit-sliced neighbour counts of the band (eval_row of schelling_bits.cuh): unsatisfied
//              mask + exact integer partials per CTA.
//   2 publish  ordered compaction of the band's unsatisfied cells into (cell, agent | type<<31)
//              records, stored straight into segment [parity][r] of EVERY rank's receive area
//              (remote stores over NVLink: the compaction IS the all-gather); CTA 0 adds the band's
//              counts; the last CTA to finish releases flag[parity][r] = step tag on every rank.
//   3 wait     one warp spins (ld.acquire.sys) on the world's flags in its OWN area, folds the counts
//              in rank order (exact integers -> the metrics row is identical on every rank), and
//              leaves u, m = min(u, e) and the ranks' prefix offsets for the movers.
//   4 move     every rank walks ALL movers k < m (keyed Feistel matching, as on one GPU): the mover's
//              record is read from the gathered segments, its target slot from the replicated E, and
//              E is updated identically everywhere.  Plane bits are flipped wherever the cell lies in
//              the rank's band OR its halo rows -- so the halo rows are maintained by the movers
//              themselves and no halo exchange exists -- and cell_agent / position / moves are
//              written by the owner of the target cell only.
// Segments are double-buffered by step parity: a rank can run at most one publish ahead of a peer
// that is still reading the previous step's records.
#pragma once
#include "common.cuh"
#include "schelling.cuh"
#include "schelling_bits.cuh"

namespace jxb {

struct GridXchgHdr {
  unsigned int flag[2][kMaxPeers];        // step tag of rank p's last publish of this parity
  unsigned int cnt[2][kMaxPeers][4];      // rank p's band: #unsatisfied, #with a neighbour, numerator lo / hi
  unsigned int err;                       // a peer's flag did not arrive within the spin budget
  unsigned int pad[128 - 2 * kMaxPeers - 8 * kMaxPeers - 1];
};
static_assert(sizeof(GridXchgHdr) == 512, "receive-area header is 512 bytes");

struct GridStepInfo {
  unsigned int tag, par, u, m;
  int key_row;                            // row of the run's key table that belongs to this step
  unsigned int ticket;                    // last-CTA election of the publish kernel
  unsigned int prefix[kMaxPeers + 1];     // rank q's records are U[prefix[q] .. prefix[q+1])
};

struct GridShardDev {
  int rank, world;
  int X0, X1;                             // rows owned by this rank
  unsigned long long cap;                 // records per (parity, rank) segment, identical on all ranks
  unsigned char* peer[kMaxPeers];         // every rank's receive area as mapped here (own = local)
  unsigned char* self;                    // == peer[rank]
  GridStepInfo* info;
  BlkPart* part;                          // [blocks] per-CTA partials of the sweep
  int blocks;                             // CTAs of the sweep AND the publish kernel (same row split)
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
__device__ __forceinline__ uint2* gs_seg(unsigned char* base, const GridShardDev& gs, unsigned int par, int from) {
  return (uint2*)(base + sizeof(GridXchgHdr)) + ((size_t)par * gs.world + from) * gs.cap;
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

// ------------------------------------------------------------------------------------ 2 publish
__global__ void __launch_bounds__(kThreads) grid_shard_publish_kernel(const SchellingDev sd, const SchellingBitsDev sb,
                                                                   const GridShardDev gs, const ModelDev md) {
  constexpr int kWarps = kThreads / 32;
  __shared__ unsigned int s_u32[kWarps], s_all[kWarps], s_occ[kWarps];
  __shared__ unsigned long long s_u64[kWarps];
  __shared__ unsigned int s_prefix, s_total, s_occ_total, s_last;
  __shared__ unsigned long long s_num_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = gridDim.x, b = blockIdx.x;
  const unsigned int tag = (unsigned int)(md.ctrl->time_step + 1);
  const unsigned int par = tag & 1u;
  const int wpr = sb.wpr;
  const int nrows = gs.X1 - gs.X0;
  const int R0 = gs.X0 + (int)((long long)nrows * b / B), R1 = gs.X0 + (int)((long long)nrows * (b + 1) / B);
  const long long wbeg = (long long)R0 * wpr, wend = (long long)R1 * wpr;
  {
    unsigned int before = 0, all = 0, occ = 0;
    unsigned long long num = 0;
    for (int i = tid; i < B; i += kThreads) {
      const uint4 raw = __ldcg((const uint4*)(gs.part + i));
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
    if (lane == 0) { s_u32[warp] = before; s_all[warp] = all; s_occ[warp] = occ; s_u64[warp] = num; }
    __syncthreads();
    if (tid == 0) {
      unsigned int p = 0, a = 0, oc = 0;
      unsigned long long nm = 0;
      for (int w = 0; w < kWarps; ++w) { p += s_u32[w]; a += s_all[w]; oc += s_occ[w]; nm += s_u64[w]; }
      s_prefix = p; s_total = a; s_occ_total = oc; s_num_total = nm;
    }
    __syncthreads();
  }
  if (b == 0 && tid < gs.world) {            // the band's counts into every rank's header
    unsigned int* c = gs_hdr(gs_peer(gs, tid))->cnt[par][gs.rank];
    c[0] = s_total;
    c[1] = s_occ_total;
    c[2] = (unsigned int)s_num_total;
    c[3] = (unsigned int)(s_num_total >> 32);
  }
  if (s_total > 0) {
    uint2* seg[kMaxPeers];
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p) seg[p] = p < gs.world ? gs_seg(gs.peer[p], gs, par, gs.rank) : nullptr;
    unsigned int base = s_prefix;
    for (long long w0 = wbeg; w0 < wend; w0 += kThreads) {
      const long long w = w0 + tid;
      unsigned int unsat = 0, tw = 0;
      if (w < wend) { unsat = __ldcg(sb.umask + w); tw = __ldcg(sb.t1 + w); }
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
      for (int ww = 0; ww < kWarps; ++ww) {
        if (ww < warp) woff += s_u32[ww];
        ttot += s_u32[ww];
      }
      unsigned int pu = base + woff + inc - cnt;
      const unsigned int c0 = (unsigned int)(w << 5);
      while (unsat) {
        const int q = __ffs(unsat) - 1;
        unsat &= unsat - 1;
        const unsigned int cell = c0 + q;
        const unsigned int ag = (unsigned int)__ldcg(sd.cell_agent + cell);
        const uint2 rec = make_uint2(cell, (ag & 0x7FFFFFFFu) | (((tw >> q) & 1u) << 31));
#pragma unroll
        for (int p = 0; p < kMaxPeers; ++p)
          if (p < gs.world) seg[p][pu] = rec;
        ++pu;
      }
      base += ttot;
    }
  }
  // last CTA to finish publishes the flag: records + counts of ALL CTAs precede it
  __threadfence_system();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&gs.info->ticket, 1u) == (unsigned int)(B - 1)) ? 1u : 0u;
  __syncthreads();
  if (s_last) {
    if (tid == 0) gs.info->ticket = 0u;
    __threadfence_system();
    if (tid < gs.world) st_release_sys(&gs_hdr(gs_peer(gs, tid))->flag[par][gs.rank], tag);
  }
}

// ------------------------------------------------------------------------------------ 3 wait
__global__ void __launch_bounds__(32) grid_shard_wait_kernel(const SchellingDev sd, const GridShardDev gs, const ModelDev md) {
  const int lane = threadIdx.x;
  Ctrl* ctrl = md.ctrl;
  const TypeDev& t = md.t[0];
  const unsigned int tag = (unsigned int)(ctrl->time_step + 1);
  const unsigned int par = tag & 1u;
  GridXchgHdr* h = gs_hdr(gs.self);
  unsigned int c0 = 0, c1 = 0;
  unsigned long long num = 0;
  if (lane < gs.world) {
    if (!*(volatile unsigned int*)&h->err) {
      const long long t0 = clock64();
      while (ld_acquire_sys(&h->flag[par][lane]) != tag) {
        if (clock64() - t0 > (20ll << 30)) { h->err = 1u; break; }     // ~10 s: a peer is gone
      }
    }
    c0 = __ldcv(&h->cnt[par][lane][0]);
    c1 = __ldcv(&h->cnt[par][lane][1]);
    num = ((unsigned long long)__ldcv(&h->cnt[par][lane][3]) << 32) | __ldcv(&h->cnt[par][lane][2]);
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

// ------------------------------------------------------------------------------------ 4 move
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

template <int MODE>
__global__ void __launch_bounds__(kThreads) grid_shard_move_kernel(const SchellingDev sd, const SchellingBitsDev sb,
                                                                const GridShardDev gs, const ModelDev md) {
  __shared__ unsigned int s_rk[8];
  __shared__ unsigned int s_prefix[kMaxPeers + 1];
  const GridStepInfo* info = gs.info;
  const unsigned int u = info->u, m = info->m;
  if (m == 0) return;
  const int tid = threadIdx.x;
  if (tid < 8) {
    const uint32_t* kp = md.keys + (size_t)info->key_row * (md.n_types + 1) * 2;
    const Key ck = {kp[0], kp[1]};
    s_rk[tid] = bits_elem<MODE>(ck, tid, 8);
  }
  if (tid <= gs.world) s_prefix[tid] = info->prefix[tid];
  __syncthreads();
  const TypeDev& t = md.t[0];
  const unsigned int e = sd.n_empty;
  const int H = sd.H, world = gs.world;
  const uint2* segs = gs_seg(gs.self, gs, info->par, 0);
  const Feistel fu = make_feistel(u, s_rk), fe = make_feistel(e, s_rk + 4);
  constexpr int kMv = 4;
  const unsigned int stride = gridDim.x * kThreads;
  for (unsigned int k0 = blockIdx.x * kThreads + tid; k0 < m; k0 += stride * kMv) {
    uint2 rec[kMv];
    unsigned int jj[kMv], dst[kMv];
    bool ok[kMv];
#pragma unroll
    for (int i = 0; i < kMv; ++i) {
      const unsigned int k = k0 + i * stride;
      ok[i] = k < m;
      rec[i] = make_uint2(0u, 0u); jj[i] = 0; dst[i] = 0;
      if (ok[i]) {
        const unsigned int j = feistel_permute(fu, k);
        int q = 0;
        while (q + 1 < world && j >= s_prefix[q + 1]) ++q;
        rec[i] = __ldcg(segs + (size_t)q * gs.cap + (j - s_prefix[q]));
        jj[i] = feistel_permute(fe, k);
        dst[i] = __ldcg(sd.E + jj[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < kMv; ++i) {
      if (!ok[i]) continue;
      const unsigned int s_ = rec[i].x, d_ = dst[i];
      const int a = (int)(rec[i].y & 0x7FFFFFFFu);
      const bool ty = (rec[i].y >> 31) != 0;
      const int xs = (int)(s_ / (unsigned int)H), xd = (int)(d_ / (unsigned int)H);
      sd.E[jj[i]] = s_;                               // replicated: identical on every rank
      gs_flip(sd, sb, gs, s_, xs, ty, false);
      gs_flip(sd, sb, gs, d_, xd, ty, true);
      if (xs >= gs.X0 && xs < gs.X1) sd.cell_agent[s_] = -1;
      if (xd >= gs.X0 && xd < gs.X1) {
        sd.cell_agent[d_] = a;
        ((int2*)t.f[1])[a] = make_int2(xd, (int)(d_ - (unsigned int)xd * (unsigned int)H));
        atomicAdd((int*)t.f[3] + a, 1);   // fire-and-forget L2 reduction: no load to wait for (ncu: 37 % of the mover stalls)
      }
    }
  }
}

// ------------------------------------------------------------------------------------ setup / export
// 'moves' is a per-agent sum over ranks: every rank counts the moves INTO its band, and the value
// uploaded by the caller stays only with the rank whose band holds the agent initially
__global__ void grid_shard_own_moves_kernel(const GridShardDev gs, const int2* pos, int* moves, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = pos[i].x;
    if (x < gs.X0 || x >= gs.X1) moves[i] = 0;
  }
}

// 'position' of the agents of the band from the cell binning (others keep the -1 fill: the host
// combines the ranks with max)
__global__ void grid_shard_export_position_kernel(const SchellingDev sd, const GridShardDev gs, int2* pos) {
  const long long c_begin = (long long)gs.X0 * sd.H, c_end = (long long)gs.X1 * sd.H;
  for (long long c = c_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; c < c_end;
       c += (long long)gridDim.x * blockDim.x) {
    const int a = sd.cell_agent[c];
    if (a >= 0) pos[a] = make_int2((int)(c / sd.H), (int)(c % sd.H));
  }
}

// 'satisfied' of the last step: the band's unsatisfied agents (movers and stayers alike) are the
// records this rank published; everybody else keeps the 1 fill (the host combines with min)
__global__ void grid_shard_export_satisfied_kernel(const GridShardDev gs, unsigned char* sat) {
  const GridStepInfo* info = gs.info;
  const unsigned int par = info->par;
  const unsigned int n = info->prefix[gs.rank + 1] - info->prefix[gs.rank];
  const uint2* seg = gs_seg(gs.self, gs, par, gs.rank);
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    sat[seg[i].y & 0x7FFFFFFFu] = 0;
}

}  // namespace jxb
