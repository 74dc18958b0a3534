// ensemble.cuh -- replica-parallel execution of whole runs (C5): the loop bodies of
// SensitivityAnalysis.run (jaxabm/analysis.py:113-157) and
// ModelCalibrator._evaluate_params_robust (jaxabm/analysis.py:434-476).
//
// One thread-block cluster (1, 2, 4 or 8 CTAs) owns one replica at a time and runs its ENTIRE time
// loop -- init from PRNGKey(seed), `steps` x (agent updates -> reduction -> env/metrics tail) --
// without leaving its SMs.  The replica's struct-of-arrays state is sliced by agent index over the
// CTAs of the cluster and lives in their shared memories (<= 200 KB per CTA); per step every CTA
// reduces its slice, stores its partial row into every peer through distributed shared memory,
// one cluster barrier, and every CTA folds the rows in rank order and runs the env/metrics tail
// on its own (identical) env copy.  Only when even 8 CTAs cannot hold the state does it fall back
// to a per-CTA scratch slot sized to stay L2-resident.  HBM sees only the final metrics rows.
#pragma once
#include <algorithm>

#include "common.cuh"
#include "rules.cuh"

namespace jxb {

constexpr int kEnsMaxSwept = 8;
constexpr int kEnsThreads = 1024;
constexpr size_t kEnsSmemBudget = 200 * 1024;

struct EnsDev {
  int program, n_types, has_env_fn, n_env, n_metrics;
  TypeDev t[JXB_MAX_TYPES];           // f[] hold byte offsets into the replica's state block
  double mp[JXB_MAX_PARAMS];
  double env0[kMaxEnv];
  int R, steps, n_swept;
  int slots[kEnsMaxSwept];            // <100: model param index; >=100: 100 + type*16 + index
  const double* params;               // [R][n_swept]
  const uint32_t* seeds;              // [R]
  double* out;                        // [R][kMaxMetrics]
  unsigned char* scratch;
  size_t state_bytes;                 // bytes of one CTA's state block
  int use_smem;
  int cluster;                        // CTAs per replica (1 when the L2 scratch path is used)
  long long slice[JXB_MAX_TYPES];     // agents per CTA of each collection (multiple of 4)
};

__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f64(double* local, unsigned int rank, double v) {
  unsigned int addr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"((unsigned int)__cvta_generic_to_shared(local)), "r"(rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(kEnsThreads) ensemble_kernel(const EnsDev ed) {
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ ModelDev md;
  __shared__ double env[kMaxEnv];
  __shared__ double s_red[(kEnsThreads / 32) * kAcc];
  __shared__ double s_tot[kAcc];
  __shared__ double s_peer[2][kMaxPeers][kAcc];     // partial rows of the cluster's CTAs, by step parity
  __shared__ double metrics[kMaxMetrics];
  const int tid = threadIdx.x;
  const int cs = ed.cluster;
  unsigned int crank = 0;
  if (cs > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int cid = blockIdx.x / cs, nclusters = gridDim.x / cs;
  unsigned char* base = ed.use_smem ? dsm : ed.scratch + (size_t)blockIdx.x * ed.state_bytes;
  unsigned int parity = 0;
  const unsigned int amask = program_acc_mask(ed.program);
  // no store into a peer's shared memory before every CTA of the cluster has started executing (compute-sanitizer
  // racecheck: "block that might not have entered yet")
  if (cs > 1) cluster_barrier();

  for (int r = cid; r < ed.R; r += nclusters) {
    if (tid == 0) {
      md.n_types = ed.n_types;
      md.program = ed.program;
      md.collect_interval = 1;
      md.has_env_fn = ed.has_env_fn;
      md.world_size = 1;
      for (int k = 0; k < JXB_MAX_PARAMS; ++k) md.mp[k] = ed.mp[k];
      for (int i = 0; i < ed.n_types; ++i) {
        md.t[i] = ed.t[i];
        for (int f = 0; f < kMaxFields; ++f) md.t[i].f[f] = base + (size_t)ed.t[i].f[f];
        md.t[i].block_begin = 0;
        md.t[i].block_count = 1;
        // this CTA's slice of the collection (the whole of it when cs == 1); per-agent keys and the
        // means keep using the global index / population (goff, gn)
        const long long lo = min((long long)crank * ed.slice[i], ed.t[i].n);
        const long long hi = min(lo + ed.slice[i], ed.t[i].n);
        md.t[i].goff = ed.t[i].goff + lo;
        md.t[i].n = hi - lo;
      }
      for (int s = 0; s < ed.n_swept; ++s) {
        const double v = ed.params[(size_t)r * ed.n_swept + s];
        const int slot = ed.slots[s];
        if (slot < 100) md.mp[slot] = v;
        else md.t[(slot - 100) / 16].p[(slot - 100) % 16] = (float)v;
      }
      for (int k = 0; k < kMaxEnv; ++k) env[k] = ed.env0[k];
      for (int k = 0; k < kMaxMetrics; ++k) metrics[k] = 0.0;
    }
    __syncthreads();
    // Model.initialize (model.py:118-144): keys = split(PRNGKey(seed), C+1)
    const Key root{0u, ed.seeds[r]};
    for (int ti = 0; ti < ed.n_types; ++ti) {
      const Key ik = split_child<MODE>(root, (unsigned long long)(ti + 1), (unsigned long long)(ed.n_types + 1));
      for (long long i = tid; i < md.t[ti].n; i += blockDim.x) init_agent<MODE>(md.t[ti], ik, i);
    }
    __syncthreads();
    for (int step = 0; step < ed.steps; ++step) {
      Acc acc;
      acc.clear();
      for (int ti = 0; ti < ed.n_types; ++ti) {
        const TypeDev& t = md.t[ti];
        switch (t.rule) {
          case JXB_RULE_RANDOM_WALKER:
          case JXB_RULE_SCALED_WALKER: rule_walker<false>(t, env, 0, acc); break;
          case JXB_RULE_CONSUMER: rule_consumer<false>(t, env, 0, acc); break;
          case JXB_RULE_PRODUCER: rule_producer<false>(t, env, 0, acc); break;
          case JXB_RULE_GROWTH: rule_growth<false>(t, 0, acc); break;
          case JXB_RULE_INCREMENT: rule_increment<false>(t, env, 0, acc); break;
          case JXB_RULE_WEALTH: rule_wealth<false>(t, 0, acc); break;
          default: break;
        }
      }
      block_reduce_acc(acc, s_red, s_tot, amask);
      __syncthreads();
      if (cs > 1) {
        // my row into every CTA of the cluster (DSMEM), one barrier, fold in rank order
        if (tid < kAcc) {
          const double v = s_tot[tid];
          for (int p = 0; p < cs; ++p) st_cluster_f64(&s_peer[parity][crank][tid], (unsigned int)p, v);
        }
        cluster_barrier();
        if (tid < kAcc) {
          const bool is_max = (tid >= kFSum && tid < kFSum + kFMax);
          double v = s_peer[parity][0][tid];
          for (int p = 1; p < cs; ++p) v = is_max ? fmax(v, s_peer[parity][p][tid]) : v + s_peer[parity][p][tid];
          s_tot[tid] = v;
        }
        parity ^= 1u;
        __syncthreads();
      }
      if (tid == 0) program_tail(md, s_tot, env, metrics);   // model.py:182-200
      __syncthreads();
    }
    if (crank == 0 && tid < kMaxMetrics) ed.out[(size_t)r * kMaxMetrics + tid] = metrics[tid];
    __syncthreads();
  }
  if (cs > 1) cluster_barrier();      // nobody leaves while a peer may still store into its shared memory
}

}  // namespace jxb
