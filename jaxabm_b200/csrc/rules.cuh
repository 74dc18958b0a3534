// rules.cuh -- fused per-type agent update rules, the env/metrics tails of the well-mixed
// programs, and the single fused step kernel that runs them.
//
// One launch per model step: CTAs [block_begin, block_begin+block_count) of the launch
// update collection `ti` (jaxabm/agent.py:132-177 for every collection of
// jaxabm/model.py:163-172), all of them reading the PRE-step env (the snapshot of
// model.py:160); every CTA leaves a partial-reduction row; the last CTA to finish
// (ticket) folds the rows in a fixed order and runs update_state_fn + metrics_fn
// (model.py:182-213) as the kernel's tail.  No host round-trip, no second launch.
//
// Arithmetic is float32 in the reference's operation order; the library is compiled with
// -fmad=false so a*b+c is two roundings exactly as XLA-CPU/NumPy evaluate the unfused ops.
#pragma once
#include "common.cuh"
#include "economy.cuh"

namespace jxb {

// ---------------------------------------------------------------------------------------
// iteration helper: groups of kVec agents, grid-stride over the CTAs owned by the type
// ---------------------------------------------------------------------------------------
#define JXB_FOR_GROUPS(t, lb, g)                                                        \
  for (long long g = (long long)(lb) * blockDim.x + threadIdx.x, _ng = (t).n / kVec,    \
                 _gs = (long long)(t).block_count * blockDim.x;                         \
       g < _ng; g += _gs)
// tail agents (n % kVec) are handled one by one by the first threads of the type's CTA 0
#define JXB_FOR_TAIL(t, lb, i)                                                          \
  for (long long i = ((t).n / kVec) * kVec + threadIdx.x; (lb) == 0 && i < (t).n; i += blockDim.x)

// ---------------------------------------------------------------------------------------
// random walker  (examples/basic_example.py:33-69; distance metric :160-163)
// fields: 0 position f32[N,2], 1 velocity f32[N,2], 2 color i32, 3 steps_taken i32
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void walker_one(float& px, float& py, float& vx, float& vy, int& color,
                                           int& steps, float lo, float hi, Acc& acc) {
  float nx = px + vx, ny = py + vy;
  const bool xb = (nx <= lo) | (nx >= hi);
  const bool yb = (ny <= lo) | (ny >= hi);
  vx = vx * (float)(1 - 2 * (int)xb);
  vy = vy * (float)(1 - 2 * (int)yb);
  px = fminf(fmaxf(nx, lo), hi);
  py = fminf(fmaxf(ny, lo), hi);
  color = (xb | yb) ? 1 - color : color;
  steps += 1;
  const float dx = px - 0.5f, dy = py - 0.5f;
  const float d = sqrtf(dx * dx + dy * dy);
  acc.fsum[0] += d;
  acc.fmax[0] = fmaxf(acc.fmax[0], d);
}

template <bool STREAM>
__device__ __forceinline__ void rule_walker(const TypeDev& t, const double* env, int lb, Acc& acc) {
  const float lo = (float)env[0], hi = (float)env[1];
  float4* pos = (float4*)t.f[0];
  float4* vel = (float4*)t.f[1];
  int4* col = (int4*)t.f[2];
  int4* stp = (int4*)t.f[3];
  JXB_FOR_GROUPS(t, lb, g) {
    float4 p0 = ldv<STREAM>(pos + 2 * g), p1 = ldv<STREAM>(pos + 2 * g + 1);
    float4 v0 = ldv<STREAM>(vel + 2 * g), v1 = ldv<STREAM>(vel + 2 * g + 1);
    int4 c = ldv<STREAM>(col + g), s = ldv<STREAM>(stp + g);
    walker_one(p0.x, p0.y, v0.x, v0.y, c.x, s.x, lo, hi, acc);
    walker_one(p0.z, p0.w, v0.z, v0.w, c.y, s.y, lo, hi, acc);
    walker_one(p1.x, p1.y, v1.x, v1.y, c.z, s.z, lo, hi, acc);
    walker_one(p1.z, p1.w, v1.z, v1.w, c.w, s.w, lo, hi, acc);
    stv<STREAM>(pos + 2 * g, p0); stv<STREAM>(pos + 2 * g + 1, p1);
    stv<STREAM>(vel + 2 * g, v0); stv<STREAM>(vel + 2 * g + 1, v1);
    stv<STREAM>(col + g, c); stv<STREAM>(stp + g, s);
  }
  JXB_FOR_TAIL(t, lb, i) {
    float2* p = (float2*)t.f[0] + i; float2* v = (float2*)t.f[1] + i;
    int* c = (int*)t.f[2] + i; int* s = (int*)t.f[3] + i;
    float2 pp = *p, vv = *v; int cc = *c, ss = *s;
    walker_one(pp.x, pp.y, vv.x, vv.y, cc, ss, lo, hi, acc);
    *p = pp; *v = vv; *c = cc; *s = ss;
  }
}

// ---------------------------------------------------------------------------------------
// consumer (tests/integration/test_integration.py:43-67)
// fields: 0 savings, 1 consumption, 2 utility, 3 income ; params: 0 base_income, 1 ptc
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void consumer_one(float& sav, float& con, float& uti, float inc, float ptc,
                                             float price, Acc& acc) {
  con = ptc * inc / price;
  sav = sav + (inc - con * price);
  uti = logf(con + 1.0f);
  acc.fsum[0] += con;
  acc.fsum[2] += uti;
}

template <bool STREAM>
__device__ __forceinline__ void rule_consumer(const TypeDev& t, const double* env, int lb, Acc& acc) {
  const float price = (float)env[0], ptc = t.p[1];
  float4* sav = (float4*)t.f[0]; float4* con = (float4*)t.f[1];
  float4* uti = (float4*)t.f[2]; const float4* inc = (const float4*)t.f[3];
  JXB_FOR_GROUPS(t, lb, g) {
    float4 s = ldv<STREAM>(sav + g), in = ldv<STREAM>(inc + g), c, u;
    consumer_one(s.x, c.x, u.x, in.x, ptc, price, acc);
    consumer_one(s.y, c.y, u.y, in.y, ptc, price, acc);
    consumer_one(s.z, c.z, u.z, in.z, ptc, price, acc);
    consumer_one(s.w, c.w, u.w, in.w, ptc, price, acc);
    stv<STREAM>(sav + g, s); stv<STREAM>(con + g, c); stv<STREAM>(uti + g, u);
  }
  JXB_FOR_TAIL(t, lb, i) {
    float s = ((float*)t.f[0])[i], c, u;
    consumer_one(s, c, u, ((float*)t.f[3])[i], ptc, price, acc);
    ((float*)t.f[0])[i] = s; ((float*)t.f[1])[i] = c; ((float*)t.f[2])[i] = u;
  }
}

// ---------------------------------------------------------------------------------------
// producer (tests/integration/test_integration.py:94-121)
// fields: 0 capital, 1 production, 2 profit ; params: 0 initial_capital, 1 productivity, 2 reinvest
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void producer_one(float& cap, float& prod, float& prof, float prd, float rr,
                                             float price, Acc& acc) {
  prod = prd * powf(cap, 0.7f);
  const float revenue = prod * price;
  const float costs = 0.1f * cap + 0.05f * prod;
  prof = revenue - costs;
  cap = cap + rr * prof;
  acc.fsum[1] += prod;
  acc.fsum[3] += prof;
}

template <bool STREAM>
__device__ __forceinline__ void rule_producer(const TypeDev& t, const double* env, int lb, Acc& acc) {
  const float price = (float)env[0], prd = t.p[1], rr = t.p[2];
  float4* cap = (float4*)t.f[0]; float4* pro = (float4*)t.f[1]; float4* prf = (float4*)t.f[2];
  JXB_FOR_GROUPS(t, lb, g) {
    float4 c = ldv<STREAM>(cap + g), p, f;
    producer_one(c.x, p.x, f.x, prd, rr, price, acc);
    producer_one(c.y, p.y, f.y, prd, rr, price, acc);
    producer_one(c.z, p.z, f.z, prd, rr, price, acc);
    producer_one(c.w, p.w, f.w, prd, rr, price, acc);
    stv<STREAM>(cap + g, c); stv<STREAM>(pro + g, p); stv<STREAM>(prf + g, f);
  }
  JXB_FOR_TAIL(t, lb, i) {
    float c = ((float*)t.f[0])[i], p, f;
    producer_one(c, p, f, prd, rr, price, acc);
    ((float*)t.f[0])[i] = c; ((float*)t.f[1])[i] = p; ((float*)t.f[2])[i] = f;
  }
}

// ---------------------------------------------------------------------------------------
// growth (tests/unit/test_analysis.py:36-39): value *= fl32(1 + growth_rate)   p[0] = factor
// increment (tests/unit/test_model.py:46-48): value += env['increment']          env[1]
// wealth (tests/unit/test_agent.py:75-84): wealth += productivity * wage_rate    p[0] = wage
// ---------------------------------------------------------------------------------------
template <bool STREAM>
__device__ __forceinline__ void rule_growth(const TypeDev& t, int lb, Acc& acc) {
  const float k = t.p[0];
  float4* val = (float4*)t.f[0];
  JXB_FOR_GROUPS(t, lb, g) {
    float4 v = ldv<STREAM>(val + g);
    v.x *= k; v.y *= k; v.z *= k; v.w *= k;
    acc.fsum[0] += v.x; acc.fsum[0] += v.y; acc.fsum[0] += v.z; acc.fsum[0] += v.w;
    stv<STREAM>(val + g, v);
  }
  JXB_FOR_TAIL(t, lb, i) {
    float v = ((float*)t.f[0])[i] * k;
    acc.fsum[0] += v;
    ((float*)t.f[0])[i] = v;
  }
}

template <bool STREAM>
__device__ __forceinline__ void rule_increment(const TypeDev& t, const double* env, int lb, Acc& acc) {
  const float inc = (float)env[1];
  float4* val = (float4*)t.f[0];
  JXB_FOR_GROUPS(t, lb, g) {
    float4 v = ldv<STREAM>(val + g);
    v.x += inc; v.y += inc; v.z += inc; v.w += inc;
    acc.fsum[0] += v.x; acc.fsum[0] += v.y; acc.fsum[0] += v.z; acc.fsum[0] += v.w;
    stv<STREAM>(val + g, v);
  }
  JXB_FOR_TAIL(t, lb, i) {
    float v = ((float*)t.f[0])[i] + inc;
    acc.fsum[0] += v;
    ((float*)t.f[0])[i] = v;
  }
}

template <bool STREAM>
__device__ __forceinline__ void rule_wealth(const TypeDev& t, int lb, Acc& acc) {
  const float wage = t.p[0];
  float4* w = (float4*)t.f[0]; const float4* pr = (const float4*)t.f[1];
  JXB_FOR_GROUPS(t, lb, g) {
    float4 v = ldv<STREAM>(w + g), p = ldv<STREAM>(pr + g);
    v.x += p.x * wage; v.y += p.y * wage; v.z += p.z * wage; v.w += p.w * wage;
    acc.fsum[0] += v.x; acc.fsum[0] += v.y; acc.fsum[0] += v.z; acc.fsum[0] += v.w;
    stv<STREAM>(w + g, v);
  }
  JXB_FOR_TAIL(t, lb, i) {
    float v = ((float*)t.f[0])[i] + ((float*)t.f[1])[i] * wage;
    acc.fsum[0] += v;
    ((float*)t.f[0])[i] = v;
  }
}

// ---------------------------------------------------------------------------------------
// env + metrics tails (update_state_fn / metrics_fn of each registered program)
// tot[]: kFSum float sums, kFMax maxima, kISum int sums, all as double
// ---------------------------------------------------------------------------------------
__device__ inline void program_tail(const ModelDev& md, const double* tot, double* env, double* m) {
  switch (md.program) {
    case JXB_PROGRAM_RANDOM_WALK: {
      // examples/basic_example.py:141-182.  The facade's overlay (agentpy.py:919-922)
      // restores Environment.state every step, so env is unchanged; mp[0] says whether the
      // collection is registered under the name compute_metrics looks up ('walkers').
      const bool live = md.mp[0] != 0.0;
      const float n = (float)md.t[0].gn;
      m[0] = env[3];                                   // mean_x
      m[1] = env[4];                                   // mean_y
      m[2] = live ? (double)((float)tot[0] / n) : 0.0; // mean_distance
      m[3] = live ? tot[kFSum + 0] : 0.0;              // max_distance
      m[4] = env[5];                                   // num_red
      m[5] = env[6];                                   // num_blue
      m[6] = env[2];                                   // time
      break;
    }
    case JXB_PROGRAM_MARKET: {
      // tests/integration/test_integration.py:125-160 then :163-183, float32 throughout
      const float total_c = (float)tot[0], total_p = (float)tot[1];
      const float rate = (float)md.mp[0];
      const float ratio = (total_p + 1e-8f) / (total_c + 1e-8f);
      const float change = rate * (1.0f - ratio);
      float price = (float)env[0] * (1.0f + change);
      price = fminf(fmaxf(price, 0.5f), 2.0f);
      const float gdp = total_p * price;
      const float unemp = fmaxf(0.0f, fminf(0.5f, 1.0f - ratio));
      env[0] = price; env[1] = gdp; env[2] = unemp; env[3] = total_c; env[4] = total_p;
      m[0] = gdp; m[1] = price; m[2] = unemp;
      // slot of each collection by rule
      for (int i = 0; i < md.n_types; ++i) {
        if (md.t[i].rule == JXB_RULE_CONSUMER) m[3] = (double)((float)tot[2] / (float)md.t[i].gn);
        if (md.t[i].rule == JXB_RULE_PRODUCER) m[4] = (double)((float)tot[3] / (float)md.t[i].gn);
      }
      break;
    }
    case JXB_PROGRAM_GROWTH: {
      // tests/unit/test_analysis.py:105-128.  env and params are Python floats there, so
      // the price recursion is float64; only jnp.mean / jnp.abs results are float32.
      const double adj = md.mp[0], target = md.mp[1];
      const double price = env[0] + adj * (target - env[0]);
      env[0] = price;
      m[0] = (double)((float)tot[0] / (float)md.t[0].gn);   // avg_value
      m[1] = price;                                          // price_level
      m[2] = (double)fabsf((float)(price - target));         // price_gap
      break;
    }
    case JXB_PROGRAM_COUNTER: {
      // tests/unit/test_model.py:20-40
      env[0] = env[0] + 1.0;
      m[0] = (double)(float)tot[0];                          // total_value
      m[1] = env[0];                                         // step_counter
      break;
    }
    default: break;
  }
}

// bookkeeping shared by every program's tail: history row + counters (model.py:203-213)
__device__ inline void finish_step(const ModelDev& md, const double* tot) {
  Ctrl* c = md.ctrl;
  const long long t = c->time_step + 1;
  // metrics are formed in a local row and copied out with plain global stores (selecting
  // between a global row and a local scratch pointer made nvcc 12.9 infer the local
  // address space for the tail's stores and drop them)
  double m[kMaxMetrics];
#pragma unroll
  for (int i = 0; i < kMaxMetrics; ++i) m[i] = 0.0;
  program_tail(md, tot, md.env, m);
  if ((t % md.collect_interval) == 0) {
    double* row = md.metrics + (size_t)c->n_recorded * kMaxMetrics;
#pragma unroll
    for (int i = 0; i < kMaxMetrics; ++i) row[i] = m[i];
    md.record_steps[c->n_recorded] = (int)t;
    c->n_recorded += 1;
  }
  c->time_step = t;
  c->step_in_run += 1;
}

// fold partial rows of all CTAs in a fixed order (warp w owns slots w, w+8, ...)
__device__ inline void fold_partials(const double* partials, int nblocks, double* tot /*smem*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = warp; i < kAcc; i += nw) {
    const bool is_max = (i >= kFSum && i < kFSum + kFMax);
    double r = is_max ? -1.0 / 0.0 : 0.0;
    for (int b = lane; b < nblocks; b += 32) {
      double v = __ldcg(partials + (size_t)b * kAcc + i);
      r = is_max ? fmax(r, v) : r + v;
    }
    r = is_max ? warp_max(r) : warp_sum(r);
    if (lane == 0) tot[i] = r;
  }
  __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(kThreads) step_kernel(const ModelDev md) {
  __shared__ double s_red[(kThreads / 32) * kAcc];
  __shared__ double s_tot[kAcc];
  __shared__ int s_last;
  int ti = 0;
  for (int i = 1; i < md.n_types; ++i)
    if ((int)blockIdx.x >= md.t[i].block_begin) ti = i;
  const TypeDev& t = md.t[ti];
  const int lb = blockIdx.x - t.block_begin;
  Acc acc;
  acc.clear();
  switch (t.rule) {
    case JXB_RULE_RANDOM_WALKER:
    case JXB_RULE_SCALED_WALKER: rule_walker<true>(t, md.env, lb, acc); break;
    case JXB_RULE_CONSUMER: rule_consumer<true>(t, md.env, lb, acc); break;
    case JXB_RULE_PRODUCER: rule_producer<true>(t, md.env, lb, acc); break;
    case JXB_RULE_GROWTH: rule_growth<true>(t, lb, acc); break;
    case JXB_RULE_INCREMENT: rule_increment<true>(t, md.env, lb, acc); break;
    case JXB_RULE_WEALTH: rule_wealth<true>(t, lb, acc); break;
    default: break;
  }
  block_reduce_acc(acc, s_red, s_tot);
  if (threadIdx.x < kAcc) md.partials[(size_t)blockIdx.x * kAcc + threadIdx.x] = s_tot[threadIdx.x];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&md.ctrl->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  fold_partials(md.partials, gridDim.x, s_tot);
  if (md.world_size > 1 && md.exchange == 1) {
    peer_exchange(md, s_tot);               // leaves the world totals in s_tot on every rank
  }
  if (threadIdx.x == 0) {
    md.ctrl->ticket = 0;
    if (md.world_size > 1 && md.exchange != 1) {
      // NCCL path: leave local totals for the all-reduce; the tail runs in tail_kernel
      for (int i = 0; i < kAcc; ++i) md.allreduce_buf[i] = s_tot[i];
    } else {
      finish_step(md, s_tot);
    }
  }
}

// tail after the cross-rank all-reduce of the partial sums (population sharding only)
__global__ void tail_kernel(const ModelDev md) {
  if (threadIdx.x == 0 && blockIdx.x == 0) finish_step(md, md.allreduce_buf);
}

// one collection, caller's key, no env/metrics: AgentCollection.update (agent.py:132-177)
template <int MODE>
__global__ void __launch_bounds__(kThreads) collection_update_kernel(const ModelDev md, int ti) {
  TypeDev t = md.t[ti];
  t.block_begin = 0;
  t.block_count = gridDim.x;
  Acc acc;
  acc.clear();
  switch (t.rule) {
    case JXB_RULE_RANDOM_WALKER:
    case JXB_RULE_SCALED_WALKER: rule_walker<true>(t, md.env, blockIdx.x, acc); break;
    case JXB_RULE_CONSUMER: rule_consumer<true>(t, md.env, blockIdx.x, acc); break;
    case JXB_RULE_PRODUCER: rule_producer<true>(t, md.env, blockIdx.x, acc); break;
    case JXB_RULE_GROWTH: rule_growth<true>(t, blockIdx.x, acc); break;
    case JXB_RULE_INCREMENT: rule_increment<true>(t, md.env, blockIdx.x, acc); break;
    case JXB_RULE_WEALTH: rule_wealth<true>(t, blockIdx.x, acc); break;
    default: break;
  }
}

// ---------------------------------------------------------------------------------------
// AgentCollection.init (agent.py:92-130): agent_keys = split(key, N); vmap(init_state)
// ---------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void init_agent(const TypeDev& t, Key key, long long i) {
  {
    const unsigned long long gi = (unsigned long long)(t.goff + i);
    switch (t.rule) {
      case JXB_RULE_RANDOM_WALKER: {  // basic_example.py:23-31 (key ignored, agentpy.py:175)
        ((float2*)t.f[0])[i] = make_float2(0.5f, 0.5f);
        ((float2*)t.f[1])[i] = make_float2(0.01f, 0.01f);
        ((int*)t.f[2])[i] = 0;
        ((int*)t.f[3])[i] = 0;
        break;
      }
      case JXB_RULE_SCALED_WALKER: {  // keyed start/velocity for the roofline run
        Key ak = split_child<MODE>(key, gi, (unsigned long long)t.gn);
        Key k0 = split_child<MODE>(ak, 0, 2), k1 = split_child<MODE>(ak, 1, 2);
        float px = bits_to_uniform(bits_scalar<MODE>(k0), 0.f, 1.f);
        float py = bits_to_uniform(bits_scalar<MODE>(k1), 0.f, 1.f);
        Key k2 = split_child<MODE>(k1, 0, 2), k3 = split_child<MODE>(k1, 1, 2);
        float vx = 0.02f * bits_to_uniform(bits_scalar<MODE>(k2), 0.f, 1.f) - 0.01f;
        float vy = 0.02f * bits_to_uniform(bits_scalar<MODE>(k3), 0.f, 1.f) - 0.01f;
        ((float2*)t.f[0])[i] = make_float2(px, py);
        ((float2*)t.f[1])[i] = make_float2(vx, vy);
        ((int*)t.f[2])[i] = 0;
        ((int*)t.f[3])[i] = 0;
        break;
      }
      case JXB_RULE_CONSUMER: {  // test_integration.py:33-41
        Key ak = split_child<MODE>(key, gi, (unsigned long long)t.gn);
        float u = bits_to_uniform(bits_scalar<MODE>(ak), 0.f, 1.f);
        ((float*)t.f[0])[i] = 0.f; ((float*)t.f[1])[i] = 0.f; ((float*)t.f[2])[i] = 0.f;
        ((float*)t.f[3])[i] = t.p[0] * (0.8f + 0.4f * u);
        break;
      }
      case JXB_RULE_PRODUCER: {  // test_integration.py:85-92
        Key ak = split_child<MODE>(key, gi, (unsigned long long)t.gn);
        float u = bits_to_uniform(bits_scalar<MODE>(ak), 0.f, 1.f);
        ((float*)t.f[0])[i] = t.p[0] * (0.8f + 0.4f * u);
        ((float*)t.f[1])[i] = 0.f; ((float*)t.f[2])[i] = 0.f;
        break;
      }
      case JXB_RULE_GROWTH: ((float*)t.f[0])[i] = t.p[1]; break;     // test_analysis.py:33-34
      case JXB_RULE_INCREMENT: {  // test_model.py:44-45
        Key ak = split_child<MODE>(key, gi, (unsigned long long)t.gn);
        ((float*)t.f[0])[i] = bits_to_uniform(bits_scalar<MODE>(ak), 0.f, 10.f);
        break;
      }
      case JXB_RULE_WEALTH: {  // test_agent.py:57-62
        Key ak = split_child<MODE>(key, gi, (unsigned long long)t.gn);
        Key k1 = split_child<MODE>(ak, 0, 2), k2 = split_child<MODE>(ak, 1, 2);
        ((float*)t.f[0])[i] = bits_to_uniform(bits_scalar<MODE>(k1), 0.f, 100.f);
        ((float*)t.f[1])[i] = bits_to_uniform(bits_scalar<MODE>(k2), 0.5f, 1.5f);
        break;
      }
      case JXB_RULE_SCHELLING: {  // schelling_model.py:26-31 (broadcast defaults)
        ((int*)t.f[0])[i] = 0;
        ((int2*)t.f[1])[i] = make_int2(0, 0);
        ((unsigned char*)t.f[2])[i] = 0;
        ((int*)t.f[3])[i] = 0;
        break;
      }
      case JXB_RULE_SIR: {  // DESIGN.md "SIR rule": I with probability p[2]
        Key ak = split_child<MODE>(key, gi, (unsigned long long)t.gn);
        float u = bits_to_uniform(bits_scalar<MODE>(ak), 0.f, 1.f);
        ((int*)t.f[0])[i] = (u < t.p[2]) ? 1 : 0;
        break;
      }
      case JXB_RULE_HOUSEHOLD: {  // advanced_economic_model.py:84-136
        const Key ak = split_child<MODE>(key, gi, (unsigned long long)t.gn);
        const float n1 = normal_scalar<MODE>(split_child<MODE>(ak, 0, 4));
        const float savings = t.p[0] * expf(n1 * 0.5f);
        const float n2 = normal_scalar<MODE>(split_child<MODE>(ak, 1, 4));
        const float income_factor = jmax(0.3f, n2 * 0.2f + 1.0f);
        const float consume_adj = beta52<MODE>(split_child<MODE>(ak, 2, 4)) * 0.4f + 0.6f;
        const bool employed = bits_to_uniform(bits_scalar<MODE>(split_child<MODE>(ak, 3, 4)), 0.f, 1.f) < 0.95f;
        ((float*)t.f[0])[i] = savings;
        ((float*)t.f[1])[i] = t.p[1] * income_factor;
        ((float*)t.f[2])[i] = savings * 0.7f;
        ((float*)t.f[3])[i] = savings * 0.3f;
        ((float*)t.f[4])[i] = 0.f;
        ((float*)t.f[5])[i] = t.p[2] * consume_adj;
        ((float*)t.f[6])[i] = t.p[3];
        ((float*)t.f[7])[i] = t.p[5];
        ((unsigned char*)t.f[8])[i] = employed ? 1 : 0;
        ((float*)t.f[9])[i] = t.p[4] * income_factor;
        ((float*)t.f[10])[i] = employed ? 1.0f : 0.0f;
        for (int f = 11; f <= 14; ++f) ((float*)t.f[f])[i] = 0.f;
        break;
      }
      case JXB_RULE_CONSUMER_FIRM: {  // advanced_economic_model.py:329-387
        const Key ak = split_child<MODE>(key, gi, (unsigned long long)t.gn);
        const float capital = t.p[0] * expf(normal_scalar<MODE>(split_child<MODE>(ak, 0, 3)) * 0.5f);
        const float eff = t.p[2] * jmax(0.5f, normal_scalar<MODE>(split_child<MODE>(ak, 1, 3)) * 0.2f + 1.0f);
        const float markup = t.p[6] * (beta52<MODE>(split_child<MODE>(ak, 2, 3)) * 0.3f + 0.1f);
        ((float*)t.f[0])[i] = capital;
        ((float*)t.f[1])[i] = eff * powf(capital, t.p[4]);
        ((float*)t.f[2])[i] = 0.f;
        ((float*)t.f[3])[i] = t.p[1];
        ((float*)t.f[4])[i] = 0.f; ((float*)t.f[5])[i] = 0.f; ((float*)t.f[6])[i] = 0.f;
        ((float*)t.f[7])[i] = eff;
        for (int f = 8; f <= 11; ++f) ((float*)t.f[f])[i] = 0.f;
        ((float*)t.f[12])[i] = 1.0f;
        ((float*)t.f[13])[i] = markup;
        ((float*)t.f[14])[i] = t.p[3]; ((float*)t.f[15])[i] = t.p[4]; ((float*)t.f[16])[i] = t.p[5];
        ((int*)t.f[17])[i] = 0;
        ((unsigned char*)t.f[18])[i] = 1;
        break;
      }
      default: break;
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads) init_kernel(const TypeDev t, Key key) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < t.n;
       i += (long long)gridDim.x * blockDim.x)
    init_agent<MODE>(t, key, i);
}

// vmap broadcast of one unbatched value of `bytes` bytes to every agent (agent.py:125-130)
__global__ void fill_kernel(unsigned char* dst, long long n, int bytes, uint4 v) {
  const unsigned char* src = (const unsigned char*)&v;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * bytes;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i % bytes];
}

}  // namespace jxb
