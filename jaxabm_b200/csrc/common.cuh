// common.cuh -- device-side model view, control block and reduction helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/jxb.h"
#include "prng.cuh"

namespace jxb {

constexpr int kMaxFields = 20;
constexpr int kMaxEnv = 32;
constexpr int kMaxMetrics = 32;
constexpr int kThreads = 256;   // CTA size of the streaming kernels
constexpr int kVec = 4;         // agents per thread per iteration (one float4 / int4 per field)

// accumulator layout of one step (what the env/metrics tail consumes)
constexpr int kFSum = 6;        // float sums  (promoted to double across blocks)
constexpr int kFMax = 2;        // float maxima
constexpr int kISum = 4;        // integer sums (exact)
constexpr int kAcc = kFSum + kFMax + kISum;

struct Acc {
  float fsum[kFSum];
  float fmax[kFMax];
  int isum[kISum];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < kFSum; ++i) fsum[i] = 0.f;
#pragma unroll
    for (int i = 0; i < kFMax; ++i) fmax[i] = -__int_as_float(0x7f800000);
#pragma unroll
    for (int i = 0; i < kISum; ++i) isum[i] = 0;
  }
};

// device-resident control block: everything that changes from step to step lives here so
// that one captured CUDA graph of a step can be replayed with no host-side arguments.
struct Ctrl {
  int step_in_run;        // index into the per-run key table
  int n_recorded;         // history rows written this run
  long long time_step;    // Model._time_step (persists across run() calls, model.py:203)
  unsigned int ticket;    // last-block election for the fused step kernel
  // Schelling per-step scalars
  unsigned int n_unsat, n_moved;
  long long total_moves;
  // SIR per-step counts
  long long sir_count[3];
  // SIR direction choice for the NEXT step (0 pull over susceptible rows, 1 push from infected rows)
  int sir_mode;             // direction of the step being executed (latched by sir_begin_step_kernel)
  int sir_mode_next;        // written by the step's tail
  long long sir_deg[2];     // adjacency entries of the susceptible / infected rows after the step
};

// cross-rank exchange of the per-step env partial sums (population sharding, one process per
// GPU): every rank owns one XchgBuf in its HBM and maps its peers' buffers through CUDA IPC
// (NVLink / NVSwitch peer memory).  The last CTA of a rank's step kernel stores its totals row
// into slot [parity][rank] of EVERY rank's buffer, publishes a sequence flag, waits for the
// world_size flags in its own buffer and folds the rows in rank order -- one kernel does the
// update, the reduction and the exchange; all ranks fold identical values in identical order.
constexpr int kMaxPeers = 8;
constexpr int kXchgRow = 20;        // doubles per exchanged row (>= kAcc and >= the economy's 15 sums + its bin range)
// a rank's exchange allocation is [XchgBuf | pad to kXchgHistOffset | income histogram u32[2^22]]: the sharded
// economy's Gini needs the GLOBAL income histogram, which every rank sums from its peers' local histograms
// straight out of this region (csrc/economy.cuh::gini_gather_kernel)
constexpr size_t kXchgHistOffset = 4096;
struct XchgBuf {
  double row[2][kMaxPeers][kXchgRow];
  unsigned int flag[2][kMaxPeers];
  unsigned int seq;       // sharded steps completed by this rank since the peers were attached
  unsigned int err;       // set when a peer's flag did not arrive within the spin budget
};

struct TypeDev {
  void* f[kMaxFields];
  long long n;            // local agents
  long long goff;         // global index of local agent 0
  long long gn;           // global population (split(key, gn))
  float p[JXB_MAX_PARAMS];
  int rule;
  int block_begin;        // first CTA of this type inside the fused launch
  int block_count;
};

struct ModelDev {
  TypeDev t[JXB_MAX_TYPES];
  int n_types;
  int program;
  int collect_interval;
  int has_env_fn;
  double* env;            // [kMaxEnv]
  double mp[JXB_MAX_PARAMS];
  Ctrl* ctrl;
  const uint32_t* keys;   // [steps][n_types+1][2] : coll keys then the update key
  double* partials;       // [grid][kAcc]
  double* metrics;        // [records][kMaxMetrics]
  int* record_steps;      // [records]
  int grid_blocks;
  int world_size;
  double* allreduce_buf;  // [kAcc] staging for the cross-rank partial-sum exchange (NCCL path)
  int rank;
  int exchange;           // 0 none, 1 in-kernel peer-memory exchange, 2 NCCL all-reduce + tail kernel
  XchgBuf* xpeer[kMaxPeers];   // every rank's buffer as mapped into this process (own = local)
  const double* consts;   // traced models: the float constants of the user code (values, not code: one
                          // compiled kernel serves every parameter set of a sweep)
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-level reduction of an Acc into `out[kAcc]` doubles (valid in threads < kAcc).
// Fixed shuffle tree + fixed warp order: bit-reproducible for a fixed launch shape.  `mask` selects
// the accumulator slots the program actually consumes (the others are left at 0 / -inf): in the
// ensemble kernel this reduction runs once per step, and a warp shuffle costs a cycle per warp.
__device__ __forceinline__ void block_reduce_acc(const Acc& a, double* smem /*[warps][kAcc]*/,
                                                 double* out, unsigned int mask = 0xFFFFFFFFu) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < kFSum; ++i) {
    if (!((mask >> i) & 1u)) continue;
    float v = warp_sum(a.fsum[i]);
    if (lane == 0) smem[warp * kAcc + i] = (double)v;
  }
#pragma unroll
  for (int i = 0; i < kFMax; ++i) {
    if (!((mask >> (kFSum + i)) & 1u)) continue;
    float v = warp_max(a.fmax[i]);
    if (lane == 0) smem[warp * kAcc + kFSum + i] = (double)v;
  }
#pragma unroll
  for (int i = 0; i < kISum; ++i) {
    if (!((mask >> (kFSum + kFMax + i)) & 1u)) continue;
    int v = warp_sum(a.isum[i]);
    if (lane == 0) smem[warp * kAcc + kFSum + kFMax + i] = (double)v;
  }
  __syncthreads();
  if (threadIdx.x < kAcc) {
    const int i = threadIdx.x;
    const bool is_max = (i >= kFSum && i < kFSum + kFMax);
    double r = is_max ? -1.0 / 0.0 : 0.0;
    if ((mask >> i) & 1u) {
      r = smem[i];
      for (int w = 1; w < nw; ++w) {
        double v = smem[w * kAcc + i];
        r = is_max ? fmax(r, v) : r + v;
      }
    }
    out[i] = r;
  }
}

// accumulator slots consumed by each program's env/metrics tail (csrc/rules.cuh::program_tail)
__device__ __host__ __forceinline__ unsigned int program_acc_mask(int program) {
  switch (program) {
    case JXB_PROGRAM_RANDOM_WALK: return 1u | (1u << kFSum);        // sum and max of the distances
    case JXB_PROGRAM_MARKET: return 0xFu;                            // consumption, production, utility, profit
    case JXB_PROGRAM_GROWTH: return 1u;
    case JXB_PROGRAM_COUNTER: return 1u;
    default: return 0u;
  }
}

// streaming load/store helpers (read-once data: bypass L1 allocation)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ int4 ld_stream(const int4* p) {
  int4 r;
  asm volatile("ld.global.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream(int4* p, int4 v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// STREAM=true: HBM-resident state, read once per step (no L1 allocation);
// STREAM=false: generic loads/stores (shared-memory or L2-resident replica state)
template <bool STREAM, class V>
__device__ __forceinline__ V ldv(const V* p) {
  if (STREAM) return ld_stream(p);
  return *p;
}
template <bool STREAM, class V>
__device__ __forceinline__ void stv(V* p, V v) {
  if (STREAM) st_stream(p, v);
  else *p = v;
}

// ---------------------------------------------------------------------------------------
// in-kernel all-reduce of the totals row over NVLink peer memory (see XchgBuf, common.cuh).
// Called by every thread of the LAST CTA of a rank's step kernel; tot = this rank's totals
// (shared memory) on entry, the world totals on exit.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int* xchg_hist(XchgBuf* b) { return (unsigned int*)((unsigned char*)b + kXchgHistOffset); }

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// tot[0..count): this rank's row on entry, the world totals on exit; entries [max_lo, max_hi) are
// combined with max, the others with +.
__device__ inline void peer_exchange(const ModelDev& md, double* tot /*smem*/, int count = kAcc, int max_lo = kFSum,
                                     int max_hi = kFSum + kFMax) {
  const int W = md.world_size, me = md.rank, tid = threadIdx.x;
  XchgBuf* mine = md.xpeer[me];
  const unsigned int seq = mine->seq + 1u;          // same value on every rank (SPMD exchange count)
  const int par = (int)(seq & 1u);
  // publish: warp p stores my row into rank p's buffer (remote stores), then the flag (release)
  const int warp = tid >> 5, lane = tid & 31;
  for (int p = warp; p < W; p += (int)(blockDim.x >> 5)) {
    XchgBuf* dst = md.xpeer[p];
    if (lane < count) dst->row[par][me][lane] = tot[lane];
    __threadfence_system();
    __syncwarp();
    if (lane == 0) st_release_sys(&dst->flag[par][me], seq);
  }
  __syncthreads();
  // wait for every rank's row of this exchange in MY buffer
  if (tid < W) {
    const long long t0 = clock64();
    while (ld_acquire_sys(&mine->flag[par][tid]) != seq) {
      if (clock64() - t0 > (20ll << 30)) { mine->err = 1u; break; }    // ~10 s at 2 GHz: a peer is gone
    }
  }
  __syncthreads();
  if (tid < count) {
    const bool is_max = (tid >= max_lo && tid < max_hi);
    double r = __ldcv(&mine->row[par][0][tid]);
    for (int p = 1; p < W; ++p) {
      const double v = __ldcv(&mine->row[par][p][tid]);
      r = is_max ? fmax(r, v) : r + v;
    }
    tot[tid] = r;
  }
  __syncthreads();
  if (tid == 0) mine->seq = seq;
  __syncthreads();
}

}  // namespace jxb
