// sir.cuh -- SIR epidemic on a Network (C3): CSR neighbour aggregation as a segmented
// reduction over ballot words, fused with the per-agent threefry draw and transition.
//
// Layout in HBM (DESIGN.md "SIR"):
//   row_ptr uint32[N+1], col int32[nnz]   adjacency binned by source from env
//                                         'network_edges' (jaxabm/agentpy.py:557,574-582)
//   state8  int8[2][N]                    S/I/R, double-buffered (pre-step snapshot
//                                         semantics of jaxabm/model.py:160)
//   infbits uint32[2][ceil(N/32)]         "is infected" bitmap the gathers hit (L2-resident)
//   rb      int32[nrb+1]                  row blocks: 32-row aligned, <= kSirTile adjacency
//                                         entries unless a single 32-row group exceeds it
//
// One CTA per row block.  The CTA streams its adjacency entries tile by tile (coalesced),
// turns each gathered infected bit into one bit of a warp ballot, prefix-sums the ballot
// popcounts, and every row reads its infected-neighbour count as a difference of two
// ranks.  Rows then draw u = uniform(split(coll_key, N)[i]) (jaxabm/agent.py:156) and
// transition; the new infected bitmap word of each 32-row group is a plain ballot store.
#pragma once
#include "common.cuh"

namespace jxb {

constexpr int kSirTile = 2048;                 // adjacency entries per tile
constexpr int kSirWords = kSirTile / 32;       // 64 ballot words
constexpr int kSirRowsPerThread = 4;           // <= 1024 rows per row block
constexpr int kSirKCap = 4095;

struct SirDev {
  const unsigned int* row_ptr;
  const int* col;
  signed char* state8[2];
  unsigned int* infbits[2];
  const int* rb;
  int nrb;
  const float* escape;     // q[k] = fl32(q[k-1]*fl32(1-beta)), k <= kSirKCap
  long long* partials;     // [max(nrb, transition CTAs)][3] new S/I/R counts per CTA
  unsigned int* k32;       // push formulation: infected-neighbour counters (zero between steps)
  const int* heavy;        // rows with more than kSirHeavy adjacency entries
  int n_heavy;
  long long* degsum;       // [CTAs][2] adjacency entries of the new susceptible / infected rows per CTA
  int auto_mode;           // 1: the step's tail picks push or pull for the next step (direction-optimising)
  unsigned int big_len;    // pull over S rows: rows longer than this are walked by the whole warp, shorter ones by their own lane
  // node-range sharding over ranks (one process per GPU): this rank owns rows [goff, goff + n) of the global
  // population; infbits[] are GLOBAL bitmaps inside an IPC-shared receive area [SirXchgHdr | bits0 | bits1], and
  // the pull kernel stores the new words of its rows straight into every rank's copy
  int world, rank;
  unsigned int gw0;              // global bitmap word of local row group 0 (goff / 32)
  unsigned int gwords;           // my bitmap words (ceil(n / 32))
  unsigned long long bits_stride;   // bytes from bits0 to bits1 inside an area
  unsigned char* peer[kMaxPeers];   // every rank's area as mapped here (own = local)
  unsigned char* self;
};

struct SirXchgHdr {
  unsigned int flag[2][kMaxPeers];        // step tag of rank p's last published step of this parity
  unsigned int cnt[2][kMaxPeers][4];      // rank p's S / I / R counts after that step
  unsigned int err;
  unsigned int pad[128 - 2 * kMaxPeers - 8 * kMaxPeers - 1];
};
static_assert(sizeof(SirXchgHdr) == 512, "receive-area header is 512 bytes");

__device__ __forceinline__ unsigned char* sir_peer(const SirDev& sv, int p) {
  unsigned char* r = sv.peer[0];
#pragma unroll
  for (int i = 1; i < kMaxPeers; ++i)
    if (p == i) r = sv.peer[i];
  return r;
}
__device__ __forceinline__ unsigned int* sir_peer_bits(const SirDev& sv, unsigned char* base, int buf) {
  return (unsigned int*)(base + sizeof(SirXchgHdr) + (size_t)buf * sv.bits_stride);
}

// ---------------------------------------------------------------------------------------
// infected-bit gather policies.
//   GatherGlobal : bitmap read through L1/L2.  A warp-wide random gather costs one L1tex
//                  wavefront per distinct 128 B line (~1 lane / cycle / SM) -- this is what bounds
//                  the plain kernel, not HBM.
//   (A variant that sliced the bitmap over the shared memories of an 8-CTA cluster and gathered
//   through distributed shared memory measured 6x SLOWER -- 4.1 ms vs 0.64 ms per step at C3 --
//   and was removed: divergent ld.shared::cluster gathers serialise per lane.  DESIGN.md 4.3.)
// ---------------------------------------------------------------------------------------
struct GatherGlobal {
  const unsigned int* inf;
  __device__ __forceinline__ unsigned int operator()(int c) const {
    return (__ldg(inf + (c >> 5)) >> (c & 31)) & 1u;
  }
};

struct SirSmem {
  unsigned int words[kSirWords + 1];
  unsigned int pref[kSirWords + 1];
  int red[3][kThreads / 32];
};

// one row block: neighbour counts by ballot-segmented reduction, threefry draw, transition,
// new bitmap words, per-block S/I/R partial counts.
template <int MODE, class Gather>
__device__ __forceinline__ void sir_row_block(const SirDev& sv, const ModelDev& md, const Gather& gather, int rbi,
                                              int cur, const Key ck, SirSmem& sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const TypeDev& t = md.t[0];
  const int nxt = cur ^ 1;
  const int r0 = sv.rb[rbi], r1 = sv.rb[rbi + 1];
  const unsigned int e0 = sv.row_ptr[r0], e1 = sv.row_ptr[r1];

  unsigned int lo[kSirRowsPerThread], hi[kSirRowsPerThread];
  int k[kSirRowsPerThread];
#pragma unroll
  for (int i = 0; i < kSirRowsPerThread; ++i) {
    const int r = r0 + i * kThreads + tid;
    k[i] = 0;
    lo[i] = hi[i] = 0;
    if (r < r1) { lo[i] = sv.row_ptr[r]; hi[i] = sv.row_ptr[r + 1]; }
  }

  for (unsigned int ts = e0; ts < e1; ts += kSirTile) {
    // all loads of the tile first (memory-level parallelism), then the gathers, then the ballots
    int cols[kSirTile / kThreads];
#pragma unroll
    for (int j = 0; j < kSirTile / kThreads; ++j) {
      const unsigned int e = ts + j * kThreads + tid;
      cols[j] = e < e1 ? __ldcs(sv.col + e) : -1;                 // streamed once
    }
    unsigned int bits[kSirTile / kThreads];
#pragma unroll
    for (int j = 0; j < kSirTile / kThreads; ++j) bits[j] = cols[j] >= 0 ? gather(cols[j]) : 0u;
#pragma unroll
    for (int j = 0; j < kSirTile / kThreads; ++j) {
      const unsigned int w = __ballot_sync(0xffffffffu, bits[j]);
      if (lane == 0) sm.words[j * (kThreads / 32) + warp] = w;
    }
    if (tid == 0) sm.words[kSirWords] = 0;
    __syncthreads();
    if (warp == 0) {  // exclusive prefix of the 64 popcounts (+ total in slot 64)
      const unsigned int a = __popc(sm.words[2 * lane]), b = __popc(sm.words[2 * lane + 1]);
      unsigned int inc = a + b;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      sm.pref[2 * lane] = inc - a - b;
      sm.pref[2 * lane + 1] = inc - b;
      if (lane == 31) sm.pref[kSirWords] = inc;
    }
    __syncthreads();
    const unsigned int te = ts + kSirTile;
#pragma unroll
    for (int i = 0; i < kSirRowsPerThread; ++i) {
      const unsigned int a = lo[i] > ts ? lo[i] : ts, b = hi[i] < te ? hi[i] : te;
      if (b > a) {
        const unsigned int xa = a - ts, xb = b - ts;
        const unsigned int ra = sm.pref[xa >> 5] + __popc(sm.words[xa >> 5] & ((1u << (xa & 31)) - 1u));
        const unsigned int rb = sm.pref[xb >> 5] + __popc(sm.words[xb >> 5] & ((1u << (xb & 31)) - 1u));
        k[i] += (int)(rb - ra);
      }
    }
    __syncthreads();
  }

  // ---- transitions ------------------------------------------------------------------------
  const float gamma = t.p[1];
  int cS = 0, cI = 0, cR = 0;
#pragma unroll
  for (int i = 0; i < kSirRowsPerThread; ++i) {
    const int r = r0 + i * kThreads + tid;
    const bool active = r < r1 && r < t.n;
    int s = 0;
    if (active) {
      s = sv.state8[cur][r];
      const bool need = (s == 0 && k[i] > 0) || s == 1;
      if (need) {
        const Key ak = split_child<MODE>(ck, (unsigned long long)(t.goff + r), (unsigned long long)t.gn);
        const float u = bits_to_uniform(bits_scalar<MODE>(ak), 0.f, 1.f);
        if (s == 0) {
          const float p = 1.0f - __ldg(sv.escape + (k[i] < kSirKCap ? k[i] : kSirKCap));
          if (u < p) s = 1;
        } else if (u < gamma) {
          s = 2;
        }
      }
      sv.state8[nxt][r] = (signed char)s;
      cS += (s == 0); cI += (s == 1); cR += (s == 2);
    }
    const unsigned int w = __ballot_sync(0xffffffffu, active && s == 1);
    // rows of a warp-iteration form one 32-aligned group -> one whole bitmap word
    const int rg = r0 + i * kThreads + warp * 32;
    if (lane == 0 && rg < r1 && rg < t.n) sv.infbits[nxt][rg >> 5] = w;
  }
  cS = warp_sum(cS); cI = warp_sum(cI); cR = warp_sum(cR);
  if (lane == 0) { sm.red[0][warp] = cS; sm.red[1][warp] = cI; sm.red[2][warp] = cR; }
  __syncthreads();
  if (tid < 3) {
    long long v = 0;
    for (int w = 0; w < kThreads / 32; ++w) v += sm.red[tid][w];
    sv.partials[(size_t)rbi * 3 + tid] = v;
  }
  __syncthreads();
}

// last CTA: exact integer fold over the per-CTA partials + metrics row (count_S, count_I, count_R);
// with_deg: also folds the adjacency sizes of the new S / I sets and picks the cheaper direction for
// the next step (cost model in DESIGN.md 4.3: an L2 reduction per infected-row entry + the
// transition pass, against a bitmap gather per susceptible-row entry).
__device__ __forceinline__ void sir_tail(const SirDev& sv, const ModelDev& md, int nparts, bool with_deg) {
  __shared__ long long s_tot[5][kThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  Ctrl* ctrl = md.ctrl;
  long long v[5] = {0, 0, 0, 0, 0};
  for (int b = tid; b < nparts; b += kThreads) {
    v[0] += __ldcg(sv.partials + (size_t)b * 3 + 0);
    v[1] += __ldcg(sv.partials + (size_t)b * 3 + 1);
    v[2] += __ldcg(sv.partials + (size_t)b * 3 + 2);
    if (with_deg) {
      v[3] += __ldcg(sv.degsum + (size_t)b * 2 + 0);
      v[4] += __ldcg(sv.degsum + (size_t)b * 2 + 1);
    }
  }
#pragma unroll
  for (int j = 0; j < 5; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
    if (lane == 0) s_tot[j][warp] = v[j];
  }
  __syncthreads();
  if (tid == 0) {
    long long c[5] = {0, 0, 0, 0, 0};
    for (int j = 0; j < 5; ++j)
      for (int w = 0; w < kThreads / 32; ++w) c[j] += s_tot[j][w];
    ctrl->ticket = 0;
    ctrl->sir_count[0] = c[0]; ctrl->sir_count[1] = c[1]; ctrl->sir_count[2] = c[2];
    if (with_deg) {
      ctrl->sir_deg[0] = c[3]; ctrl->sir_deg[1] = c[4];
      // cost model fitted on B200 at C3, microseconds with dS, dI in millions of entries, n = 10 M:
      //   push ~ 75 + 23 min(dI, 10) + 9.5 max(dI - 10, 0) + 3.0 dS   (L2 reductions + transition pass + the
      //                                                                draws of the exposed S rows)
      //   pull ~ 145 + 2.9 dS + 1.0 dI                                 (bitmap gathers over the S rows)
      // => pull wins once the infected rows hold more than ~2 % of n*10 entries: 220 dI + dS > 50 n
      if (sv.auto_mode) {
        const long long nn = md.t[0].n, dS = c[3], dI = c[4];
        ctrl->sir_mode_next = (220 * dI + dS > 50 * nn) ? 0 : 1;
      }
    }
    const long long tsn = ctrl->time_step + 1;
    if ((tsn % md.collect_interval) == 0) {
      double* row = md.metrics + (size_t)ctrl->n_recorded * kMaxMetrics;
      row[0] = (double)c[0]; row[1] = (double)c[1]; row[2] = (double)c[2];
      md.record_steps[ctrl->n_recorded] = (int)tsn;
      ctrl->n_recorded += 1;
    }
    ctrl->time_step = tsn;
    ctrl->step_in_run += 1;
  }
}

// plain variant: one CTA per row block, bitmap gathered through L1/L2
template <int MODE>
__global__ void __launch_bounds__(kThreads) sir_step_kernel(const SirDev sv, const ModelDev md) {
  __shared__ SirSmem sm;
  __shared__ int s_last;
  Ctrl* ctrl = md.ctrl;
  const int cur = (int)(ctrl->time_step & 1);
  const uint32_t* kp = md.keys + (size_t)ctrl->step_in_run * (md.n_types + 1) * 2;
  const Key ck = {kp[0], kp[1]};
  GatherGlobal g{sv.infbits[cur]};
  sir_row_block<MODE>(sv, md, g, blockIdx.x, cur, ck, sm);
  __threadfence();
  if (threadIdx.x == 0) s_last = (atomicAdd(&ctrl->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  sir_tail(sv, md, sv.nrb, false);
}

// ---------------------------------------------------------------------------------------
// push formulation (frontier-driven): only the adjacency rows of INFECTED agents are read.
//   sir_push_kernel       for every infected row j, for every neighbour i: k32[i] += 1
//                         (fire-and-forget L2 reductions; k32 is L2-resident, 4 B / agent)
//   sir_transition_kernel per agent: k = k32[i] (re-zeroed), draw, transition, new bitmap word,
//                         S/I/R partial counts; last CTA folds + metrics row.
// Work is proportional to the infected agents' edges instead of all edges, the counts are the same
// exact integers as in the pull kernel, so the two formulations are bit-identical.
// Rows longer than kSirHeavy entries (static list built with the CSR) are spread over a whole
// CTA; all other rows are handled by one warp per 32-row group, lanes striding the row.
// ---------------------------------------------------------------------------------------
constexpr int kSirHeavy = 2048;

// first kernel of every step: latch the direction picked by the previous step's tail, so that the
// three self-gating kernels of ONE step all see the same decision
__global__ void sir_begin_step_kernel(Ctrl* ctrl) {
  if (threadIdx.x == 0) ctrl->sir_mode = ctrl->sir_mode_next;
}

// CTA-level fold of the S/I/R counts and the S/I adjacency sizes into this CTA's partial rows
__device__ __forceinline__ void sir_publish_partials(const SirDev& sv, int cS, int cI, int cR, long long dS, long long dI,
                                                     int (*s_red)[kThreads / 32]) {
  __shared__ long long s_deg[2][kThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  cS = warp_sum(cS); cI = warp_sum(cI); cR = warp_sum(cR);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dS += __shfl_xor_sync(0xffffffffu, dS, o);
    dI += __shfl_xor_sync(0xffffffffu, dI, o);
  }
  if (lane == 0) { s_red[0][warp] = cS; s_red[1][warp] = cI; s_red[2][warp] = cR; s_deg[0][warp] = dS; s_deg[1][warp] = dI; }
  __syncthreads();
  if (tid < 3) {
    long long v = 0;
    for (int w = 0; w < kThreads / 32; ++w) v += s_red[tid][w];
    sv.partials[(size_t)blockIdx.x * 3 + tid] = v;
  } else if (tid < 5) {
    long long v = 0;
    for (int w = 0; w < kThreads / 32; ++w) v += s_deg[tid - 3][w];
    sv.degsum[(size_t)blockIdx.x * 2 + (tid - 3)] = v;
  }
}

__device__ __forceinline__ void red_add_u32(unsigned int* p, unsigned int v) {
  asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(kThreads) sir_push_kernel(const SirDev sv, const ModelDev md) {
  const int tid = threadIdx.x, lane = tid & 31;
  const Ctrl* ctrl = md.ctrl;
  if (ctrl->sir_mode != 1) return;                 // this step runs in the pull direction
  const int cur = (int)(ctrl->time_step & 1);
  const unsigned int* inf = sv.infbits[cur];
  const long long n = md.t[0].n;
  // heavy rows: one CTA per row, all threads stride the row
  for (int h = blockIdx.x; h < sv.n_heavy; h += gridDim.x) {
    const int r = sv.heavy[h];
    if (!((__ldg(inf + (r >> 5)) >> (r & 31)) & 1u)) continue;
    const unsigned int lo = sv.row_ptr[r], hi = sv.row_ptr[r + 1];
    for (unsigned int e = lo + tid; e < hi; e += kThreads) red_add_u32(sv.k32 + __ldcs(sv.col + e), 1u);
  }
  // Warp w owns the groups g = w (mod #warps), as before (the hubs sit at low indices: consecutive groups must go to
  // different warps), but it fetches the bitmap words of 32 of its groups with ONE load instruction (lane l reads
  // the word of its l-th next group) and then visits the groups that hold an infected row.  (One word per
  // warp-iteration made the scan a chain of ~33 dependent L2 round trips per warp: 23 us per step even when almost
  // nobody is infected.)
  const long long ngroups = (n + 31) >> 5;
  const long long nwarps = (long long)gridDim.x * (kThreads / 32);
  const long long wid = (long long)blockIdx.x * (kThreads / 32) + (tid >> 5);
  for (long long i0 = 0; wid + i0 * nwarps < ngroups; i0 += 32) {
    const long long my_g = wid + (i0 + lane) * nwarps;
    const unsigned int my_word = my_g < ngroups ? __ldg(inf + my_g) : 0u;
    unsigned int live = __ballot_sync(0xffffffffu, my_word != 0u);
    while (live) {
      const int gl = __ffs(live) - 1;
      live &= live - 1;
      unsigned int word = __shfl_sync(0xffffffffu, my_word, gl);
      const long long rbase = (wid + (i0 + gl) * nwarps) << 5;
      const unsigned int my_lo = (rbase + lane <= n) ? sv.row_ptr[rbase + lane] : 0u;
      const unsigned int last = (rbase + 32 <= n) ? sv.row_ptr[rbase + 32] : sv.row_ptr[n];
      while (word) {
        const int b = __ffs(word) - 1;
        word &= word - 1;
        const unsigned int lo = __shfl_sync(0xffffffffu, my_lo, b);
        const unsigned int hi = b == 31 ? last : __shfl_sync(0xffffffffu, my_lo, (b + 1) & 31);
        if (hi - lo > (unsigned)kSirHeavy) continue;             // done by the heavy-row pass
        for (unsigned int e = lo + lane; e < hi; e += 32) red_add_u32(sv.k32 + __ldcs(sv.col + e), 1u);
      }
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads) sir_transition_kernel(const SirDev sv, const ModelDev md) {
  __shared__ int s_red[3][kThreads / 32];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  Ctrl* ctrl = md.ctrl;
  if (ctrl->sir_mode != 1) return;                 // the pull kernel does its own transitions
  const TypeDev& t = md.t[0];
  const int cur = (int)(ctrl->time_step & 1), nxt = cur ^ 1;
  const uint32_t* kp = md.keys + (size_t)ctrl->step_in_run * (md.n_types + 1) * 2;
  const Key ck = {kp[0], kp[1]};
  const float gamma = t.p[1];
  long long dS = 0, dI = 0;
  const signed char* __restrict__ st_cur = sv.state8[cur];
  signed char* __restrict__ st_nxt = sv.state8[nxt];
  unsigned int* __restrict__ k32 = sv.k32;
  int cS = 0, cI = 0, cR = 0;
  // persistent: a CTA sweeps tiles of kThreads*kSirRowsPerThread rows; a warp-iteration covers 32
  // consecutive rows, i.e. exactly one word of the new infected bitmap
  constexpr int kTile = kThreads * kSirRowsPerThread;
  for (long long base = (long long)blockIdx.x * kTile; base < t.n; base += (long long)gridDim.x * kTile) {
    int s[kSirRowsPerThread];
    unsigned int k[kSirRowsPerThread], deg[kSirRowsPerThread];
#pragma unroll
    for (int i = 0; i < kSirRowsPerThread; ++i) {       // all loads of the tile in flight first
      const long long r = base + i * kThreads + tid;
      s[i] = 0; k[i] = 0; deg[i] = 0;
      if (r < t.n) { s[i] = st_cur[r]; k[i] = __ldcg(k32 + r); deg[i] = sv.row_ptr[r + 1] - sv.row_ptr[r]; }
    }
#pragma unroll
    for (int i = 0; i < kSirRowsPerThread; ++i) {
      const long long r = base + i * kThreads + tid;
      const bool active = r < t.n;
      int sn = s[i];
      if (active) {
        if (k[i]) k32[r] = 0u;
        const bool need = (sn == 0 && k[i] > 0) || sn == 1;
        if (need) {
          const Key ak = split_child<MODE>(ck, (unsigned long long)(t.goff + r), (unsigned long long)t.gn);
          const float u = bits_to_uniform(bits_scalar<MODE>(ak), 0.f, 1.f);
          if (sn == 0) {
            const float p = 1.0f - __ldg(sv.escape + (k[i] < (unsigned)kSirKCap ? k[i] : (unsigned)kSirKCap));
            if (u < p) sn = 1;
          } else if (u < gamma) {
            sn = 2;
          }
        }
        st_nxt[r] = (signed char)sn;
        cS += (sn == 0); cI += (sn == 1); cR += (sn == 2);
        if (sn == 0) dS += deg[i];
        if (sn == 1) dI += deg[i];
      }
      const unsigned int w = __ballot_sync(0xffffffffu, active && sn == 1);
      const long long rg = base + i * kThreads + warp * 32;
      if (lane == 0 && rg < t.n) sv.infbits[nxt][rg >> 5] = w;
    }
  }
  sir_publish_partials(sv, cS, cI, cR, dS, dI, s_red);
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&ctrl->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  sir_tail(sv, md, (int)gridDim.x, true);
}

// ---------------------------------------------------------------------------------------
// pull over the SUSCEPTIBLE rows only (fused step, no atomics): one warp per 32-row group, lane =
// row.  Only susceptible agents need their infected-neighbour count, so only their adjacency is
// read: each lane walks its own susceptible row (long rows: all 32 lanes stride the row), one
// bitmap gather per entry.  Transitions, the new bitmap word and the partial counts follow in
// the same warp.  Cost is proportional to the susceptible rows' adjacency -- the complement of the
// push kernel's; the step's tail picks whichever is cheaper for the next step.
// ---------------------------------------------------------------------------------------
template <int MODE, bool SHARD = false>
__global__ void __launch_bounds__(kThreads) sir_pull_s_kernel(const SirDev sv, const ModelDev md) {
  __shared__ int s_red[3][kThreads / 32];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31;
  Ctrl* ctrl = md.ctrl;
  if (!SHARD && ctrl->sir_mode != 0) return;
  const TypeDev& t = md.t[0];
  const long long n = t.n;
  const int cur = (int)(ctrl->time_step & 1), nxt = cur ^ 1;
  const uint32_t* kp = md.keys + (size_t)ctrl->step_in_run * (md.n_types + 1) * 2;
  const Key ck = {kp[0], kp[1]};
  const float gamma = t.p[1];
  const unsigned int* __restrict__ inf = sv.infbits[cur];
  const signed char* __restrict__ st_cur = sv.state8[cur];
  signed char* __restrict__ st_nxt = sv.state8[nxt];
  int cS = 0, cI = 0, cR = 0;
  long long dS = 0, dI = 0;
  const long long ngroups = (n + 31) >> 5;
  const long long wstride = (long long)gridDim.x * (kThreads / 32);
  for (long long g = (long long)blockIdx.x * (kThreads / 32) + (tid >> 5); g < ngroups; g += wstride) {
    const long long r = (g << 5) + lane;
    const bool active = r < n;
    int s = 3;
    unsigned int lo = 0, len = 0;
    if (active) {
      s = st_cur[r];
      lo = sv.row_ptr[r];
      len = sv.row_ptr[r + 1] - lo;
    }
    unsigned int k = 0;
    // long susceptible rows (hubs before they are infected): all 32 lanes stride the row
    unsigned int big = __ballot_sync(0xffffffffu, active && s == 0 && len > sv.big_len);
    while (big) {
      const int b = __ffs(big) - 1;
      big &= big - 1;
      const unsigned int blo = __shfl_sync(0xffffffffu, lo, b), blen = __shfl_sync(0xffffffffu, len, b);
      unsigned int cnt = 0;
      for (unsigned int e0 = 0; e0 < blen; e0 += 128) {          // 4 independent gathers per lane in flight
        unsigned int bits = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const unsigned int e = e0 + j * 32 + lane;
          if (e < blen) {
            const int c = __ldcs(sv.col + blo + e);
            bits += (__ldg(inf + (c >> 5)) >> (c & 31)) & 1u;
          }
        }
        cnt += bits;
      }
      cnt = (unsigned int)warp_sum((int)cnt);
      if (lane == b) k = cnt;
    }
    // every other susceptible row: its own lane walks it (consecutive lanes own consecutive CSR
    // segments, so a warp's loads fall into a few adjacent lines that L1 keeps across the
    // iterations); two entries per iteration for memory-level parallelism
    if (active && s == 0 && len > 0 && len <= sv.big_len) {
      const int* cp = sv.col + lo;
      unsigned int e = 0;
      for (; e + 2 <= len; e += 2) {
        const int c0 = __ldg(cp + e), c1 = __ldg(cp + e + 1);
        const unsigned int w0 = __ldg(inf + (c0 >> 5)), w1 = __ldg(inf + (c1 >> 5));
        k += ((w0 >> (c0 & 31)) & 1u) + ((w1 >> (c1 & 31)) & 1u);
      }
      if (e < len) {
        const int c0 = __ldg(cp + e);
        k += (__ldg(inf + (c0 >> 5)) >> (c0 & 31)) & 1u;
      }
    }
    int sn = s;
    if (active) {
      const bool need = (s == 0 && k > 0) || s == 1;
      if (need) {
        const Key ak = split_child<MODE>(ck, (unsigned long long)(t.goff + r), (unsigned long long)t.gn);
        const float u = bits_to_uniform(bits_scalar<MODE>(ak), 0.f, 1.f);
        if (s == 0) {
          const float p = 1.0f - __ldg(sv.escape + (k < (unsigned)kSirKCap ? k : (unsigned)kSirKCap));
          if (u < p) sn = 1;
        } else if (u < gamma) {
          sn = 2;
        }
      }
      st_nxt[r] = (signed char)sn;
      cS += (sn == 0); cI += (sn == 1); cR += (sn == 2);
      if (sn == 0) dS += len;
      if (sn == 1) dI += len;
    }
    const unsigned int w = __ballot_sync(0xffffffffu, active && sn == 1);
    if (SHARD) {
      // the new word of this 32-row group goes straight into EVERY rank's next bitmap (remote stores over
      // NVLink): the aggregation kernel is also the all-gather of the state slices
      if (lane < sv.world) sir_peer_bits(sv, sir_peer(sv, lane), nxt)[sv.gw0 + g] = w;
    } else {
      if (lane == 0) sv.infbits[nxt][g] = w;
    }
  }
  sir_publish_partials(sv, cS, cI, cR, dS, dI, s_red);
  if (SHARD) __threadfence_system(); else __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&ctrl->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  if (!SHARD) {
    __threadfence();
    sir_tail(sv, md, (int)gridDim.x, true);
    return;
  }
  // sharded: fold this rank's exact counts, hand them to every rank and release the step's flag; the
  // wait kernel that follows folds the ranks and writes the metrics row
  __threadfence_system();
  __shared__ long long s_cnt[3];
  __shared__ long long s_w[3][kThreads / 32];
  {
    long long v[3] = {0, 0, 0};
    for (int b = tid; b < (int)gridDim.x; b += kThreads) {
#pragma unroll
      for (int j = 0; j < 3; ++j) v[j] += __ldcg(sv.partials + (size_t)b * 3 + j);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
      if (lane == 0) s_w[j][tid >> 5] = v[j];
    }
    __syncthreads();
    if (tid < 3) {
      long long c = 0;
      for (int w = 0; w < kThreads / 32; ++w) c += s_w[tid][w];
      s_cnt[tid] = c;
    }
  }
  __syncthreads();
  const unsigned int tag = (unsigned int)(ctrl->time_step + 1);
  const unsigned int par = tag & 1u;
  if (tid < sv.world) {
    SirXchgHdr* h = (SirXchgHdr*)sir_peer(sv, tid);
    h->cnt[par][sv.rank][0] = (unsigned int)s_cnt[0];
    h->cnt[par][sv.rank][1] = (unsigned int)s_cnt[1];
    h->cnt[par][sv.rank][2] = (unsigned int)s_cnt[2];
    __threadfence_system();
    st_release_sys(&h->flag[par][sv.rank], tag);
  }
  if (tid == 0) ctrl->ticket = 0;
}

// sharded step, second launch: wait for every rank's flag of this step (their bitmap slices and counts are
// then in my area), fold the counts in rank order and write the metrics row (what sir_tail does on one GPU)
__global__ void __launch_bounds__(32) sir_shard_wait_kernel(const SirDev sv, const ModelDev md) {
  const int lane = threadIdx.x;
  Ctrl* ctrl = md.ctrl;
  const unsigned int tag = (unsigned int)(ctrl->time_step + 1);
  const unsigned int par = tag & 1u;
  SirXchgHdr* h = (SirXchgHdr*)sv.self;
  long long c0 = 0, c1 = 0, c2 = 0;
  if (lane < sv.world) {
    if (!*(volatile unsigned int*)&h->err) {
      const long long t0 = clock64();
      while (ld_acquire_sys(&h->flag[par][lane]) != tag) {
        if (clock64() - t0 > (20ll << 30)) { h->err = 1u; break; }     // ~10 s: a peer is gone
      }
    }
    c0 = __ldcv(&h->cnt[par][lane][0]);
    c1 = __ldcv(&h->cnt[par][lane][1]);
    c2 = __ldcv(&h->cnt[par][lane][2]);
  }
  __syncwarp();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c0 += __shfl_xor_sync(0xffffffffu, c0, o);
    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    c2 += __shfl_xor_sync(0xffffffffu, c2, o);
  }
  if (lane == 0) {
    ctrl->sir_count[0] = c0; ctrl->sir_count[1] = c1; ctrl->sir_count[2] = c2;
    const long long tsn = ctrl->time_step + 1;
    if ((tsn % md.collect_interval) == 0) {
      double* row = md.metrics + (size_t)ctrl->n_recorded * kMaxMetrics;
      row[0] = (double)c0; row[1] = (double)c1; row[2] = (double)c2;
      md.record_steps[ctrl->n_recorded] = (int)tsn;
      ctrl->n_recorded += 1;
    }
    ctrl->time_step = tsn;
    ctrl->step_in_run += 1;
  }
}

// after the API column was (re)packed: copy my slice of the current bitmap into every peer's copy
// (setup path; the host barriers afterwards)
__global__ void sir_shard_sync_kernel(const SirDev sv, int cur) {
  const unsigned int* mine = sir_peer_bits(sv, sv.self, cur) + sv.gw0;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < sv.gwords; i += gridDim.x * blockDim.x) {
    const unsigned int w = mine[i];
    for (int p = 0; p < sv.world; ++p)
      if (p != sv.rank) sir_peer_bits(sv, sir_peer(sv, p), cur)[sv.gw0 + i] = w;
  }
}

// ---------------------------------------------------------------------------------------
// CSR construction on the device: env 'network_edges' int32[E,2] (jaxabm/agentpy.py:557) -> row_ptr /
// col by counting sort on the source: degree histogram (L2 reductions), exclusive prefix scan,
// scatter through per-row cursors.  The order of a row's neighbours is whatever the atomics make
// it; every formulation above only counts, so results do not depend on it.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) csr_count_kernel(const int2* edges, long long n_edges, long long n,
                                                             long long n_cols, unsigned int* deg, int* err) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (long long)gridDim.x * blockDim.x) {
    const int2 ed = __ldcs(edges + e);
    if (ed.x < 0 || ed.x >= n || ed.y < 0 || ed.y >= n_cols) { atomicExch(err, 1); continue; }
    atomicAdd(deg + ed.x, 1u);
  }
}

constexpr int kScanTile = kThreads * 16;

// tile sums of v[0..n)
__global__ void __launch_bounds__(kThreads) scan_tile_sums_kernel(const unsigned int* v, long long n, unsigned int* sums) {
  __shared__ unsigned int s_w[kThreads / 32];
  const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * 16;
  unsigned int s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += base + j < n ? v[base + j] : 0u;
  s = (unsigned int)warp_sum((int)s);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int tot = 0;
    for (int w = 0; w < kThreads / 32; ++w) tot += s_w[w];
    sums[blockIdx.x] = tot;
  }
}

// in-place exclusive scan of the tile sums by ONE CTA (any count), total written to sums[count]
__global__ void __launch_bounds__(1024) scan_sums_kernel(unsigned int* sums, int count) {
  __shared__ unsigned int s_w[32];
  __shared__ unsigned int s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int c0 = 0; c0 < count; c0 += 1024) {
    const int i = c0 + tid;
    const unsigned int v = i < count ? sums[i] : 0u;
    unsigned int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    unsigned int off = s_carry;
    for (int w = 0; w < warp; ++w) off += s_w[w];
    if (i < count) sums[i] = off + inc - v;
    __syncthreads();
    if (tid == 1023) s_carry = off + inc;
    __syncthreads();
  }
  if (tid == 0) sums[count] = s_carry;
}

// out[i] = exclusive prefix of v (out may alias v); also copies the prefix into `copy` when given;
// out[n] = total
__global__ void __launch_bounds__(kThreads) scan_apply_kernel(const unsigned int* v, long long n, const unsigned int* sums,
                                                              unsigned int* out, unsigned int* copy) {
  __shared__ unsigned int s_w[kThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long base = (long long)blockIdx.x * kScanTile + (long long)tid * 16;
  unsigned int c[16];
  unsigned int s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) { c[j] = base + j < n ? v[base + j] : 0u; s += c[j]; }
  unsigned int inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  unsigned int run = sums[blockIdx.x] + inc - s;
  for (int w = 0; w < warp; ++w) run += s_w[w];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (base + j < n) { out[base + j] = run; if (copy) copy[base + j] = run; }
    run += c[j];
  }
  if (blockIdx.x == gridDim.x - 1 && tid == kThreads - 1) out[n] = sums[gridDim.x];
}

// (the same range test as csr_count_kernel: an edge it flagged and skipped must not be scattered either)
__global__ void __launch_bounds__(kThreads) csr_fill_kernel(const int2* edges, long long n_edges, long long n, long long n_cols,
                                                            unsigned int* cursor, int* col) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (long long)gridDim.x * blockDim.x) {
    const int2 ed = __ldcs(edges + e);
    if (ed.x < 0 || ed.x >= n || ed.y < 0 || ed.y >= n_cols) continue;
    const unsigned int pos = atomicAdd(cursor + ed.x, 1u);
    col[pos] = ed.y;
  }
}

// API column 'state' int32[N]  <->  packed int8 + infected bitmap
__global__ void sir_pack_kernel(const int* state, signed char* s8, unsigned int* bits, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int s = 0;
  if (i < n) { s = state[i]; s8[i] = (signed char)s; }
  const unsigned int w = __ballot_sync(0xffffffffu, i < n && s == 1);
  if ((threadIdx.x & 31) == 0 && i < n) bits[i >> 5] = w;      // bits = word of local row 0 (sharded: + goff / 32)
}

__global__ void sir_unpack_kernel(const signed char* s8, int* state, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    state[i] = (int)s8[i];
}

// the same, for a snapshot taken between the launches of a step (possibly replayed from a graph): the current
// buffer is chosen by the DEVICE-side step counter
__global__ void sir_unpack_cur_kernel(const SirDev sv, const Ctrl* ctrl, int* state, long long n) {
  const signed char* s8 = sv.state8[(int)(ctrl->time_step & 1)];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    state[i] = (int)s8[i];
}

}  // namespace jxb
