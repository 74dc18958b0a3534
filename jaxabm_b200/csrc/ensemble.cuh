// ensemble.cuh -- replica-parallel execution of whole runs (C5): the loop bodies of
// SensitivityAnalysis.run (jaxabm/analysis.py:113-157) and
// ModelCalibrator._evaluate_params_robust (jaxabm/analysis.py:434-476).
//
// One CTA owns one replica at a time and runs its ENTIRE time loop -- init from
// PRNGKey(seed), `steps` x (agent updates -> block reduction -> env/metrics tail) -- without
// leaving the SM.  Replica state lives in the CTA's shared memory when it fits (<= 200 KB),
// otherwise in a per-CTA scratch slot that is sized to stay L2-resident (grid x state <=
// ~64 MB of the 126 MB L2).  HBM sees only the final metrics row of every replica.
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
  size_t state_bytes;
  int use_smem;
};

template <int MODE>
__global__ void __launch_bounds__(kEnsThreads) ensemble_kernel(const EnsDev ed) {
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ ModelDev md;
  __shared__ double env[kMaxEnv];
  __shared__ double s_red[(kEnsThreads / 32) * kAcc];
  __shared__ double s_tot[kAcc];
  __shared__ double metrics[kMaxMetrics];
  const int tid = threadIdx.x;
  unsigned char* base = ed.use_smem ? dsm : ed.scratch + (size_t)blockIdx.x * ed.state_bytes;

  for (int r = blockIdx.x; r < ed.R; r += gridDim.x) {
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
      block_reduce_acc(acc, s_red, s_tot);
      __syncthreads();
      if (tid == 0) program_tail(md, s_tot, env, metrics);   // model.py:182-200
      __syncthreads();
    }
    if (tid < kMaxMetrics) ed.out[(size_t)r * kMaxMetrics + tid] = metrics[tid];
    __syncthreads();
  }
}

}  // namespace jxb
