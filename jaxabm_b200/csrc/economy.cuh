// economy.cuh -- C4-B: households + consumer-goods firms (examples/models/advanced_economic_model.py
// of the reference, capital/energy firms and the climate/pandemic modules off).
//
//   Household.update          :138-296   -> household_one     (1 uniform of split(key,4))
//   ConsumerGoodsFirm.update  :389-581   -> firm_one          (1 normal + 1 uniform of split(key,4))
//   update_environment        :1461-1738 -> eco_update_environment (2 normals of split(update_key,5))
//   compute_metrics           :1741-1905 -> eco_compute_metrics    (29 metrics, Gini)
//
// One model step = economy_step_kernel (both collections in one launch, 14 float sums + 1 count as
// per-CTA partial rows + the 22-bit radix histogram of the new household incomes; last CTA folds the
// rows and runs update_environment) -> Gini kernels (scan of the histogram, rank-weighted sum) whose last
// CTA runs compute_metrics and appends the history row.  All of it is captured in the step graph.
//
// Arithmetic: float32 in the reference's operation order (-fmad=false); Python-float env entries
// are doubles and scalar (op) scalar expressions are evaluated in double as Python does; NaNs
// propagate through minimum/maximum/clip as in jax.numpy (CUDA's fminf/fmaxf would drop them) --
// the reference model does produce NaNs (0/0 once a firm's inventory covers its demand share), and
// compute_metrics' nan_to_num defaults are part of the observable output.
#pragma once
#include "common.cuh"

namespace jxb {

// env slots of JXB_PROGRAM_ECONOMY (csrc/engine.cu kPrograms keeps the same order)
enum {
  EE_WAGE = 0, EE_PRICE_LEVEL, EE_INTEREST, EE_GDP, EE_GDP_GROWTH, EE_LABOR_SUPPLY, EE_LABOR_DEMAND,
  EE_EMPLOYMENT, EE_UNEMPLOYMENT, EE_TIME_STEP, EE_CLIMATE, EE_PANDEMIC,          // rewritten every step
  EE_TAX_RATE, EE_ENERGY_PRICE, EE_JOB_MARKET, EE_GOODS_AVAIL, EE_CG_DEMAND, EE_CG_PRICE, EE_KG_PRICE,
  EE_INFLATION, EE_DEBT_TO_GDP, EE_AVG_UTILITY, EE_INCOME_PC, EE_GOVT_SPENDING, EE_ENERGY_SUPPLY,
  EE_KG_DEMAND,                                                                   // kept by :1731-1735
  EE_TOTAL_INCOME, EE_TOTAL_INCOME_SET,                                           // frozen at step 1
  EE_GINI,                                                                        // scratch: Gini of the step
  EE_COUNT
};

constexpr int kEcoF = 14;             // float sums
constexpr int kEcoAcc = kEcoF + 1;    // + employed count
constexpr int kEcoRow = kEcoAcc + 2;  // + the populated range of the income histogram: max bin, -(min bin) (both fold with max)
// Gini histogram: the top kGiniBits bits of the order-preserving integer image of the float = bins of 2^-10
// relative width.  Incomes sharing a bin get their mean rank; the rank-weighted sum then differs from the sorted
// one by about (bin width) / (6 * populated bins) ~ 1e-7 relative, far inside the 1e-5 tolerance.
// (Round 1 used 22 bits; the populated range then spans ~16 K bins and every household update is a global L2
// reduction -- measured bound of the whole step.  With 19 bits the populated range fits a per-CTA shared-memory
// window.)
constexpr int kGiniBits = 19;
constexpr int kGiniBins = 1 << kGiniBits;
constexpr int kGiniScanTile = 4096;   // bins per scan CTA
constexpr int kGiniWin = 8192;        // bins of the per-CTA shared-memory histogram window (8 octaves of income)

struct EcoDev {
  double* partials;          // [grid][kEcoRow]
  unsigned int* hist;        // [kGiniBins] where THIS rank's households count their incomes: bin_count itself on one
                             // GPU; the histogram region of the rank's exchange allocation when the population is sharded
  unsigned int* bin_count;   // [kGiniBins] counts of the WHOLE population (sharded: summed from the ranks' hist by gini_gather_kernel)
  unsigned int* bin_base;    // [kGiniBins] exclusive prefix of the counts
  int* tile_range;           // [0], [1]: first / last scan tile holding an income of this step (written by the step's
                             // tail); [2]: first tile of the two-tile window that held the most incomes (written by the
                             // scan, or seeded from a sample of the initial incomes) -- where the NEXT step's kernels
                             // put their shared-memory histogram window; -1 = unknown
  unsigned int* scan_sums;   // [kGiniBins / kGiniScanTile]
  double* gini_partials;     // [gini grid][2]
  unsigned int* ticket2;     // election ticket of the Gini accumulate kernel
  int gini_blocks;
};

// jax.numpy semantics: NaN in, NaN out
__device__ __forceinline__ float jmax(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }
__device__ __forceinline__ float jmin(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b); }
__device__ __forceinline__ float jclip(float x, float lo, float hi) { return jmin(jmax(x, lo), hi); }

// XLA's float32 erf_inv (Giles' polynomial), as restated in oracle/jaxlike.py::erfinv_f32
__device__ __forceinline__ float erfinv_f32(float x) {
  float w = -log1pf(-(x * x));
  float p;
  if (w < 5.0f) {
    w = w - 2.5f;
    p = 2.81022636e-08f;
    p = 3.43273939e-07f + p * w;
    p = -3.5233877e-06f + p * w;
    p = -4.39150654e-06f + p * w;
    p = 0.00021858087f + p * w;
    p = -0.00125372503f + p * w;
    p = -0.00417768164f + p * w;
    p = 0.246640727f + p * w;
    p = 1.50140941f + p * w;
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    p = 0.000100950558f + p * w;
    p = 0.00134934322f + p * w;
    p = -0.00367342844f + p * w;
    p = 0.00573950773f + p * w;
    p = -0.0076224613f + p * w;
    p = 0.00943887047f + p * w;
    p = 1.00167406f + p * w;
    p = 2.83297682f + p * w;
  }
  if (fabsf(x) == 1.0f) return copysignf(__int_as_float(0x7f800000), x);
  return p * x;
}

// jax.random.normal(key, ()): sqrt(2) * erfinv(uniform(nextafter(-1,0), 1))
template <int MODE>
__device__ __forceinline__ float normal_scalar(Key k) {
  const float lo = __int_as_float(0xbf7fffff);        // nextafter(-1, 0)
  const float u = bits_to_uniform(bits_scalar<MODE>(k), lo, 1.0f);
  return 1.41421354f * erfinv_f32(u);
}

// Beta(5,2) = G5 / (G5 + G2), G_a = -sum log1p(-u_i), u = uniform(key, (7,)).  Builder-authored
// stand-in for jax.random.beta (oracle/economy.py docstring).
template <int MODE>
__device__ __forceinline__ float beta52(Key k) {
  float e[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) e[j] = -log1pf(-bits_to_uniform(bits_elem<MODE>(k, j, 7), 0.f, 1.f));
  const float g5 = (((e[0] + e[1]) + e[2]) + e[3]) + e[4];
  const float g2 = e[5] + e[6];
  return g5 / (g5 + g2);
}

struct EcoEnvView {      // what the agent rules read, hoisted once per CTA
  float job_loss, job_find, wage, tax, transfer, price_level, goods_av, rate_h;
  float md_share, market_price, e_price, c_price, rate_f, climate, pandemic;
};

__device__ __forceinline__ EcoEnvView eco_env_view(const double* env) {
  EcoEnvView v;
  const double shock = env[EE_JOB_MARKET] * env[EE_PANDEMIC];
  v.job_loss = (float)(0.02 / shock);
  v.job_find = (float)(0.1 * shock);
  v.wage = (float)env[EE_WAGE];
  v.tax = (float)env[EE_TAX_RATE];
  v.transfer = 0.5f;                                   // 'transfer_rate' is never in env (:180)
  v.price_level = (float)env[EE_PRICE_LEVEL];
  v.goods_av = (float)env[EE_GOODS_AVAIL];
  v.rate_h = (float)env[EE_INTEREST];
  v.md_share = (float)(env[EE_CG_DEMAND] * 0.01);      // market_demand * firm_market_share (:427-430)
  v.market_price = (float)env[EE_CG_PRICE];
  v.e_price = (float)env[EE_ENERGY_PRICE];
  v.c_price = 1.0f;                                    // 'capital_price' is never in env (:417)
  v.rate_f = (float)env[EE_INTEREST];
  v.climate = (float)env[EE_CLIMATE];
  v.pandemic = (float)env[EE_PANDEMIC];
  return v;
}

// ---------------------------------------------------------------------------------------
// Household.update (:138-296).  fields: 0 savings 1 income 2 bank_deposits 3 cash 4 debt
// 5 propensity_to_consume 6 propensity_to_save 7 risk_aversion 8 employed(bool) 9 productivity
// 10 labor_supply 11 consumption 12 utility 13 taxes_paid 14 transfers_received
// params: 0 initial_savings 1 initial_income 2 ptc 3 pts 4 labor_productivity 5 risk_aversion
// ---------------------------------------------------------------------------------------
struct HouseholdIO {
  float income, deposits, cash, ptc, pts, risk, productivity;
  bool employed;
  // outputs
  float savings, labor_supply, consumption, utility, taxes, transfers;
};

__device__ __forceinline__ void household_one(HouseholdIO& h, float rv, const EcoEnvView& v, float init_inc,
                                              float two_init_inc) {
  const bool new_emp = h.employed ? (rv > v.job_loss) : (rv < v.job_find);
  const float labor_supply = new_emp ? 1.0f : 0.0f;
  const float labor_income = (labor_supply * h.productivity) * v.wage;
  const float transfers = ((1.0f - (new_emp ? 1.0f : 0.0f)) * h.income) * v.transfer;
  const float gross = labor_income + transfers;
  const float base_tax = gross * v.tax;
  const float prog = gross > two_init_inc ? 0.05f * (gross / init_inc - 2.0f) : 0.0f;
  const float taxes = base_tax * (1.0f + prog);
  const float net = gross - taxes;
  const float desired = net * h.ptc;
  const float avail = h.cash + h.deposits * 0.3f;
  const float actual = jmin(desired, avail) * v.goods_av;
  const float real_c = actual / v.price_level;
  const float interest_income = h.deposits * v.rate_h;
  const float target = net * h.pts;
  const float adj = target - h.deposits * 0.1f;
  const float dep = jmin(adj, h.cash - actual);
  const float pos = jmax(0.0f, dep), neg = jmax(0.0f, -dep);
  const float new_cash = ((h.cash - actual) - pos) + neg;
  const float new_dep = (h.deposits + pos) + interest_income;
  const float cu = log1pf(real_c);
  const float su = h.risk * log1pf(new_dep / 100.0f);
  h.savings = new_cash + new_dep;
  h.income = gross;
  h.deposits = new_dep;
  h.cash = new_cash;
  h.employed = new_emp;
  h.labor_supply = labor_supply;
  h.consumption = real_c;
  h.utility = cu + su;
  h.taxes = taxes;
  h.transfers = transfers;
}

__device__ __forceinline__ unsigned int gini_bin(float x);

// count one income: into the CTA's shared-memory window when the bin falls inside it (the window is placed on the
// previous step's populated range), straight into the global histogram otherwise
// Lanes of the warp that hit the same bin are counted by ONE atomic (match_any): shared-memory atomics on one
// address serialise lane by lane, and a degenerate income distribution -- the reference model's incomes all turn
// NaN after a few steps, i.e. one single bin -- would otherwise cost 32 serialised atomics per warp instruction
// (measured: 4.9 ms per step instead of 1.0).
__device__ __forceinline__ void hist_count(unsigned int* bin_count, unsigned int* s_hist, unsigned int win_lo, unsigned int bin) {
  const unsigned int peers = __match_any_sync(__activemask(), bin);
  if ((threadIdx.x & 31) != (unsigned int)(__ffs(peers) - 1)) return;
  const unsigned int cnt = __popc(peers), off = bin - win_lo;
  if (off < (unsigned int)kGiniWin) atomicAdd(s_hist + off, cnt);
  else atomicAdd(bin_count + bin, cnt);
}

// One agent's draw: u = uniform(split(split(coll_key, N)[i], 4)[0]) -- three dependent threefry blocks
// (agent.py:156 then advanced_economic_model.py:159,174).
template <int MODE>
__device__ __forceinline__ float household_draw(Key ck, unsigned long long gi, unsigned long long gn) {
  const Key ak = split_child<MODE>(ck, gi, gn);
  return bits_to_uniform(bits_scalar<MODE>(split_child<MODE>(ak, 0, 4)), 0.f, 1.f);
}

// Four agents per thread per iteration: one 16-byte load / store per float column (4-byte for the bool column),
// and four independent threefry / update chains in flight per thread -- the kernel is bound by instruction issue
// with long dependent chains (3 x 20 threefry rounds per agent), so instruction-level parallelism is what fills
// the issue slots that one chain per thread leaves empty.
template <int MODE>
__device__ __forceinline__ void rule_household(const TypeDev& t, const double* env, Key ck, int lb, float* fs,
                                               int& employed_count, unsigned int* bin_count, unsigned int& bin_lo,
                                               unsigned int& bin_hi, unsigned int* s_hist, unsigned int win_lo) {
  const EcoEnvView v = eco_env_view(env);
  const float init_inc = t.p[1];
  const float two_init_inc = (float)(2.0 * (double)t.p[1]);
  float* savings = (float*)t.f[0]; float* income = (float*)t.f[1]; float* deposits = (float*)t.f[2];
  float* cash = (float*)t.f[3]; const float* ptc = (const float*)t.f[5]; const float* pts = (const float*)t.f[6];
  const float* risk = (const float*)t.f[7]; unsigned char* employed = (unsigned char*)t.f[8];
  const float* productivity = (const float*)t.f[9]; float* labor_supply = (float*)t.f[10];
  float* consumption = (float*)t.f[11]; float* utility = (float*)t.f[12]; float* taxes = (float*)t.f[13];
  float* transfers = (float*)t.f[14];
  const long long stride = (long long)t.block_count * blockDim.x;
  const long long ngroups = t.n / 4;
  const unsigned long long gn = (unsigned long long)t.gn;
  for (long long g = (long long)lb * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
    const float4 f_inc = ld_stream((const float4*)income + g), f_dep = ld_stream((const float4*)deposits + g);
    const float4 f_cash = ld_stream((const float4*)cash + g), f_ptc = ld_stream((const float4*)ptc + g);
    const float4 f_pts = ld_stream((const float4*)pts + g), f_risk = ld_stream((const float4*)risk + g);
    const float4 f_prod = ld_stream((const float4*)productivity + g);
    const uchar4 f_emp = __ldcs((const uchar4*)employed + g);
    HouseholdIO h[4];
    h[0].income = f_inc.x; h[1].income = f_inc.y; h[2].income = f_inc.z; h[3].income = f_inc.w;
    h[0].deposits = f_dep.x; h[1].deposits = f_dep.y; h[2].deposits = f_dep.z; h[3].deposits = f_dep.w;
    h[0].cash = f_cash.x; h[1].cash = f_cash.y; h[2].cash = f_cash.z; h[3].cash = f_cash.w;
    h[0].ptc = f_ptc.x; h[1].ptc = f_ptc.y; h[2].ptc = f_ptc.z; h[3].ptc = f_ptc.w;
    h[0].pts = f_pts.x; h[1].pts = f_pts.y; h[2].pts = f_pts.z; h[3].pts = f_pts.w;
    h[0].risk = f_risk.x; h[1].risk = f_risk.y; h[2].risk = f_risk.z; h[3].risk = f_risk.w;
    h[0].productivity = f_prod.x; h[1].productivity = f_prod.y; h[2].productivity = f_prod.z; h[3].productivity = f_prod.w;
    h[0].employed = f_emp.x != 0; h[1].employed = f_emp.y != 0; h[2].employed = f_emp.z != 0; h[3].employed = f_emp.w != 0;
    float rv[4];
    const unsigned long long gi0 = (unsigned long long)(t.goff + 4 * g);
#pragma unroll
    for (int j = 0; j < 4; ++j) rv[j] = household_draw<MODE>(ck, gi0 + j, gn);
#pragma unroll
    for (int j = 0; j < 4; ++j) household_one(h[j], rv[j], v, init_inc, two_init_inc);
    st_stream((float4*)savings + g, make_float4(h[0].savings, h[1].savings, h[2].savings, h[3].savings));
    st_stream((float4*)income + g, make_float4(h[0].income, h[1].income, h[2].income, h[3].income));
    st_stream((float4*)deposits + g, make_float4(h[0].deposits, h[1].deposits, h[2].deposits, h[3].deposits));
    st_stream((float4*)cash + g, make_float4(h[0].cash, h[1].cash, h[2].cash, h[3].cash));
    __stcs((uchar4*)employed + g, make_uchar4(h[0].employed ? 1 : 0, h[1].employed ? 1 : 0, h[2].employed ? 1 : 0, h[3].employed ? 1 : 0));
    st_stream((float4*)labor_supply + g, make_float4(h[0].labor_supply, h[1].labor_supply, h[2].labor_supply, h[3].labor_supply));
    st_stream((float4*)consumption + g, make_float4(h[0].consumption, h[1].consumption, h[2].consumption, h[3].consumption));
    st_stream((float4*)utility + g, make_float4(h[0].utility, h[1].utility, h[2].utility, h[3].utility));
    st_stream((float4*)taxes + g, make_float4(h[0].taxes, h[1].taxes, h[2].taxes, h[3].taxes));
    st_stream((float4*)transfers + g, make_float4(h[0].transfers, h[1].transfers, h[2].transfers, h[3].transfers));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      fs[0] += h[j].labor_supply; fs[1] += h[j].consumption; fs[2] += h[j].savings; fs[3] += h[j].deposits;
      fs[4] += h[j].income; fs[5] += h[j].utility;
      employed_count += h[j].employed ? 1 : 0;
      // Gini histogram of the NEW incomes (metrics are computed on the post-update state): fire-and-forget L2
      // reductions whose latency hides behind the streaming loads
      const unsigned int bin = gini_bin(h[j].income);
      bin_lo = min(bin_lo, bin); bin_hi = max(bin_hi, bin);
      hist_count(bin_count, s_hist, win_lo, bin);
    }
  }
  // the last t.n % 4 agents
  for (long long i = ngroups * 4 + threadIdx.x; lb == 0 && i < t.n; i += blockDim.x) {
    HouseholdIO h;
    h.income = income[i]; h.deposits = deposits[i]; h.cash = cash[i];
    h.ptc = ptc[i]; h.pts = pts[i]; h.risk = risk[i];
    h.productivity = productivity[i]; h.employed = employed[i] != 0;
    const float rv = household_draw<MODE>(ck, (unsigned long long)(t.goff + i), gn);
    household_one(h, rv, v, init_inc, two_init_inc);
    savings[i] = h.savings; income[i] = h.income; deposits[i] = h.deposits;
    cash[i] = h.cash; employed[i] = (unsigned char)(h.employed ? 1 : 0);
    labor_supply[i] = h.labor_supply; consumption[i] = h.consumption;
    utility[i] = h.utility; taxes[i] = h.taxes; transfers[i] = h.transfers;
    fs[0] += h.labor_supply; fs[1] += h.consumption; fs[2] += h.savings; fs[3] += h.deposits;
    fs[4] += h.income; fs[5] += h.utility;
    employed_count += h.employed ? 1 : 0;
    const unsigned int bin = gini_bin(h.income);
    bin_lo = min(bin_lo, bin); bin_hi = max(bin_hi, bin);
    hist_count(bin_count, s_hist, win_lo, bin);
  }
}

// ---------------------------------------------------------------------------------------
// ConsumerGoodsFirm.update (:389-581).  fields: 0 capital_stock 1 production_capacity 2 inventory
// 3 cash 4 revenue 5 profit 6 debt 7 production_efficiency 8 labor_demand 9 energy_usage
// 10 goods_produced 11 goods_sold 12 price 13 markup_rate 14 labor_elasticity 15 capital_elasticity
// 16 energy_elasticity 17 age(i32) 18 is_active(bool)
// params: 0 initial_capital 1 initial_cash 2 production_efficiency 3 labor_el 4 capital_el 5 energy_el 6 markup
// ---------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void rule_firm(const TypeDev& t, const double* env, Key ck, int lb, float* fs) {
  const EcoEnvView v = eco_env_view(env);
  float* capital = (float*)t.f[0]; float* capacity = (float*)t.f[1]; float* inventory = (float*)t.f[2];
  float* cash = (float*)t.f[3]; float* revenue_c = (float*)t.f[4]; float* profit_c = (float*)t.f[5];
  const float* debt = (const float*)t.f[6]; const float* efficiency = (const float*)t.f[7];
  float* labor_c = (float*)t.f[8]; float* energy_c = (float*)t.f[9]; float* produced_c = (float*)t.f[10];
  float* sold_c = (float*)t.f[11]; float* price_c = (float*)t.f[12]; const float* markup = (const float*)t.f[13];
  const float* lel = (const float*)t.f[14]; const float* cel = (const float*)t.f[15];
  const float* eel = (const float*)t.f[16]; int* age = (int*)t.f[17]; unsigned char* active = (unsigned char*)t.f[18];
  const long long stride = (long long)t.block_count * blockDim.x;
  for (long long i = (long long)lb * blockDim.x + threadIdx.x; i < t.n; i += stride) {
    const float K = capital[i], inv = inventory[i], csh = cash[i], dbt = debt[i], eff = efficiency[i];
    const float price = price_c[i], mk = markup[i], le = lel[i], ce = cel[i], ee = eel[i];
    const Key ak = split_child<MODE>(ck, (unsigned long long)(t.goff + i), (unsigned long long)t.gn);
    const float tp = jmax(0.0f, v.md_share - inv);
    const float Kc = powf(K, ce);
    const float denom1 = (eff * Kc) * powf(1.0f, ee);
    const float base_labor = powf(tp / denom1, 1.0f / le);
    const float max_labor = csh / v.wage;
    const float labor = jmin(base_labor, max_labor);
    const float Ll = powf(labor, le);
    const float denom2 = (eff * Kc) * Ll;
    const float base_energy = powf(tp / denom2, 1.0f / ee);
    const float remaining = csh - labor * v.wage;
    const float max_energy = remaining / v.e_price;
    const float energy = jmin(base_energy, max_energy);
    const float production = ((eff * Ll) * Kc) * powf(energy, ee);
    const float affected = (production * v.climate) * v.pandemic;
    const float noise = normal_scalar<MODE>(split_child<MODE>(ak, 0, 4)) * 0.05f + 1.0f;
    const float actual_prod = affected * noise;
    const float new_inv = inv + actual_prod;
    const float cost = (labor * v.wage + energy * v.e_price) + (K * v.c_price) * 0.05f;
    const float one_mk = 1.0f + mk;
    const float unit_cost = actual_prod > 0.0f ? cost / actual_prod : price / one_mk;
    const float target_price = unit_cost * one_mk;
    const float new_price = price * (float)(1 - 0.3) + target_price * 0.3f;
    const float pressure = inv / (actual_prod + 0.1f);
    const float discount = jmax(0.0f, jmin(0.2f, 0.05f * pressure));
    const float final_price = new_price * (1.0f - discount);
    const float competitiveness = powf(v.market_price / final_price, 1.2f);
    const float sales_rand = bits_to_uniform(bits_scalar<MODE>(split_child<MODE>(ak, 1, 4)), 0.f, 1.f) * 0.4f + 0.8f;
    const float potential = (v.md_share * competitiveness) * sales_rand;
    const float sales = jmin(potential, new_inv);
    const float final_inv = new_inv - sales;
    const float revenue = sales * final_price;
    const float total_costs = cost + dbt * v.rate_f;
    const float profit = revenue - total_costs;
    const float new_cash = ((csh + revenue) - labor * v.wage) - energy * v.e_price;
    const float inv_ratio = profit > 0.0f ? 0.3f : 0.0f;
    const float investment = profit * inv_ratio;
    const float actual_inv = jmin(investment, new_cash * 0.5f);
    const float purchases = actual_inv / v.c_price;
    const float new_K = K * (float)(1 - 0.05) + purchases;
    const float final_cash = new_cash - actual_inv;
    const bool viable = (final_cash > 0.0f) & (new_K > 0.0f);
    capital[i] = new_K; capacity[i] = new_K * eff; inventory[i] = final_inv; cash[i] = final_cash;
    revenue_c[i] = revenue; profit_c[i] = profit; labor_c[i] = labor; energy_c[i] = energy;
    produced_c[i] = actual_prod; sold_c[i] = sales; price_c[i] = final_price;
    age[i] = age[i] + 1; active[i] = viable ? 1 : 0;
    fs[6] += actual_prod; fs[7] += sales; fs[8] += final_inv; fs[9] += final_price; fs[10] += labor;
    fs[11] += energy; fs[12] += profit; fs[13] += new_K;
  }
}

// ---------------------------------------------------------------------------------------
// update_environment (:1461-1738).  tot[0..13] float sums (as doubles), tot[14] employed count.
// Only the entries outside the carry-over list (:1731-1735) change; total_income freezes at its
// first value.
// ---------------------------------------------------------------------------------------
template <int MODE>
__device__ inline void eco_update_environment(const ModelDev& md, const double* tot, double* env, Key update_key) {
  int hi = -1, fi = -1;
  for (int i = 0; i < md.n_types; ++i) {
    if (md.t[i].rule == JXB_RULE_HOUSEHOLD) hi = i;
    if (md.t[i].rule == JXB_RULE_CONSUMER_FIRM) fi = i;
  }
  if (hi < 0) return;                                          // "No households": env unchanged (:1483-1484)
  const float n_h = (float)md.t[hi].gn;
  const float total_labor_supply = (float)tot[0];
  const float total_income = (float)tot[4];
  const float employment_rate = (float)tot[14] / n_h;
  float produced = 0.f, cg_price = 1.0f, cf_labor = 0.f;
  if (fi >= 0) {
    produced = (float)tot[6];
    cg_price = (float)tot[9] / (float)md.t[fi].gn;
    cf_labor = (float)tot[10];
  }
  const float old_wage = (float)env[EE_WAGE], old_price = (float)env[EE_PRICE_LEVEL], old_rate = (float)env[EE_INTEREST];
  const float total_labor_demand = (cf_labor + 0.0f) + 0.0f;
  const float tightness = total_labor_supply > 1e-6f ? total_labor_demand / (total_labor_supply + 1e-6f) : 1.0f;
  const float wage_pressure = (tightness - 1.0f) * 0.2f;
  const float wage_noise = normal_scalar<MODE>(split_child<MODE>(update_key, 0, 5)) * 0.01f;
  const float wage_change = jclip(wage_pressure + wage_noise, -0.05f, 0.05f);
  const float new_wage = jmax(0.1f, old_wage * (1.0f + wage_change));
  const float p0 = 0.6f * cg_price, p1 = 0.3f * 2.0f, p2 = 0.1f * 1.0f;
  const float new_price = jmax(0.1f, (p0 + p1) + p2);
  const float inflation = new_price / jmax(0.1f, old_price) - 1.0f;
  const float unemployment = 1.0f - employment_rate;
  const float inflation_gap = inflation - 0.02f;
  const float output_gap = -0.5f * (unemployment - 0.05f);
  const float taylor = (0.02f + 1.5f * inflation_gap) + 0.5f * output_gap;
  float change = (taylor - old_rate) * 0.3f;
  const float rate_noise = normal_scalar<MODE>(split_child<MODE>(update_key, 1, 5)) * 0.005f;
  change = jclip(change + rate_noise, -0.02f, 0.02f);
  const float new_rate = jmax(0.01f, old_rate + change);
  float gdp = produced * cg_price + (float)(0.0 * 2.0);
  gdp = gdp + 0.0f;
  gdp = jmax(0.1f, gdp);
  const float prev_gdp = jmax(0.1f, (float)env[EE_GDP]);
  const float gdp_growth = gdp / prev_gdp - 1.0f;
  env[EE_TIME_STEP] = env[EE_TIME_STEP] + 1.0;
  env[EE_WAGE] = new_wage;
  env[EE_LABOR_SUPPLY] = total_labor_supply;
  env[EE_LABOR_DEMAND] = total_labor_demand;
  env[EE_EMPLOYMENT] = employment_rate;
  env[EE_UNEMPLOYMENT] = 1.0f - employment_rate;
  env[EE_PRICE_LEVEL] = new_price;
  env[EE_INTEREST] = new_rate;
  env[EE_GDP] = gdp;
  env[EE_GDP_GROWTH] = gdp_growth;
  env[EE_CLIMATE] = 1.0;
  env[EE_PANDEMIC] = 1.0;
  if (env[EE_TOTAL_INCOME_SET] == 0.0) { env[EE_TOTAL_INCOME] = total_income; env[EE_TOTAL_INCOME_SET] = 1.0; }
}

// compute_metrics (:1741-1905).  m[29] in the order of kPrograms' metric list.
__device__ inline float nn(float v, float dflt) { return v != v ? dflt : v; }

__device__ inline void eco_compute_metrics(const double* env, float gini, double* m) {
  const float gdp = (float)env[EE_GDP], gdp_growth = (float)env[EE_GDP_GROWTH];
  const float inflation = (float)env[EE_INFLATION], unemployment = (float)env[EE_UNEMPLOYMENT];
  const float wage = (float)env[EE_WAGE], tightness = (float)env[EE_JOB_MARKET];
  const float goods_av = (float)env[EE_GOODS_AVAIL], c_price = (float)env[EE_CG_PRICE];
  const float k_price = (float)env[EE_KG_PRICE], e_price = (float)env[EE_ENERGY_PRICE];
  const float rate = (float)env[EE_INTEREST], debt_gdp = (float)env[EE_DEBT_TO_GDP];
  const float avg_utility = (float)env[EE_AVG_UTILITY], ipc = (float)env[EE_INCOME_PC];
  const double tech = 1.0, renew = 0.2, carbon = 0.0;
  const double climate = env[EE_CLIMATE], pandemic = env[EE_PANDEMIC];
  // sector GDPs: *_sold entries are never written to env -> 0.0 * price (Python floats)
  const double g_gdp = env[EE_GOVT_SPENDING];
  const double total = 0.0 + 0.0 + 0.0 + g_gdp;
  const float govt_share = total > 0.1 ? (float)((g_gdp / total) * 100.0) : 0.0f;
  float health = 0.25f * (1.0f - unemployment);
  health = health + 0.15f * jclip(gdp_growth * 10.0f, -1.0f, 1.0f);
  health = health + 0.15f * (1.0f - fabsf(inflation - 0.02f) * 10.0f);
  health = health + 0.10f * (1.0f - jmin(debt_gdp, 1.0f));
  health = health + 0.10f * goods_av;
  health = health + (0.10f * avg_utility) / 2.0f;
  health = health + 0.05f * (1.0f - jmin(gini, 1.0f));
  health = health + (0.05f * (float)tech) / 2.0f;
  health = health + 0.05f * (float)renew;
  const float health_index = jclip(health * 100.0f, 0.0f, 100.0f);
  const double tclip = fmin(fmax((tech - 1.0) * 0.5, 0.0), 1.0);
  const double sustain = (0.4 * renew + 0.3 * (1.0 - fmin(carbon / 100.0, 1.0)) + 0.2 * tclip +
                          0.1 * (1.0 - fmax(0.0, climate - 1.0))) * 100.0;
  m[0] = nn(gdp, 0.1f); m[1] = nn(gdp_growth * 100.0f, 0.0f); m[2] = nn(inflation * 100.0f, 0.0f);
  m[3] = nn(unemployment * 100.0f, 0.0f); m[4] = nn(wage, 1.0f); m[5] = nn(rate * 100.0f, 0.0f);
  m[6] = nn(goods_av * 100.0f, 100.0f); m[7] = nn(tightness, 1.0f); m[8] = nn(c_price, 1.0f);
  m[9] = nn(k_price, 2.0f); m[10] = nn(e_price, 1.0f); m[11] = nn(avg_utility, 0.0f); m[12] = nn(ipc, 0.0f);
  m[13] = nn(gini, 0.0f); m[14] = 0.0; m[15] = 0.0; m[16] = 0.0; m[17] = nn(govt_share, 0.0f);
  m[18] = (float)tech; m[19] = nn((float)env[EE_KG_DEMAND], 0.0f); m[20] = 0.0;
  m[21] = nn((float)env[EE_ENERGY_SUPPLY], 0.0f); m[22] = (float)(renew * 100.0); m[23] = (float)carbon;
  m[24] = nn((float)sustain, 50.0f); m[25] = nn(debt_gdp * 100.0f, 0.0f);
  m[26] = nn((float)((1.0 - (climate - 1.0)) * 100.0), 0.0f); m[27] = nn((float)((1.0 - pandemic) * 100.0), 0.0f);
  m[28] = nn(health_index, 50.0f);
}

// ---------------------------------------------------------------------------------------
// step kernel: both collections in one launch; last CTA folds partial rows in a fixed order and
// runs update_environment.  (Metrics + history row follow in gini_accumulate_kernel.)
// ---------------------------------------------------------------------------------------
// One launch per collection (RULE is a template parameter so that the household path is not
// compiled with the firm path's register footprint); the launches of a step share the election
// ticket: the last CTA of the LAST launch folds all `total_ctas` partial rows.
template <int MODE, int RULE>
__global__ void __launch_bounds__(kThreads, 3) economy_step_kernel(const ModelDev md, const EcoDev ed, int ti,
                                                                int row_offset, int total_ctas) {
  __shared__ double s_red[(kThreads / 32) * kEcoRow];
  __shared__ double s_tot[kEcoRow];
  __shared__ int s_last;
  __shared__ unsigned int s_hist[RULE == JXB_RULE_HOUSEHOLD ? kGiniWin : 1];
  const TypeDev& t = md.t[ti];
  const int lb = blockIdx.x;
  // shared-memory histogram window = the two scan tiles (8 octaves of income) that held the most incomes in the
  // previous step: the bulk of the distribution moves by a few percent per step, while its tails (long-term
  // unemployed whose transfers halve every step) spread over many octaves and stay on global reductions
  unsigned int win_lo = 0xFFFFFFFFu - (unsigned int)kGiniWin;      // disabled: every bin is "outside"
  if (RULE == JXB_RULE_HOUSEHOLD) {
    const int wt = ed.tile_range[2];
    if (wt >= 0) win_lo = (unsigned int)wt * kGiniScanTile;
    for (int i = threadIdx.x; i < kGiniWin; i += kThreads) s_hist[i] = 0u;
    __syncthreads();
  }
  const uint32_t* kp = md.keys + (size_t)md.ctrl->step_in_run * (md.n_types + 1) * 2;
  const Key ck = {kp[2 * ti], kp[2 * ti + 1]};
  float fs[kEcoF];
#pragma unroll
  for (int i = 0; i < kEcoF; ++i) fs[i] = 0.f;
  int emp = 0;
  unsigned int bin_lo = kGiniBins, bin_hi = 0;        // empty range: hi < lo
  if (RULE == JXB_RULE_HOUSEHOLD) rule_household<MODE>(t, md.env, ck, lb, fs, emp, ed.hist, bin_lo, bin_hi, s_hist, win_lo);
  else rule_firm<MODE>(t, md.env, ck, lb, fs);
  if (RULE == JXB_RULE_HOUSEHOLD) {
    __syncthreads();
    for (int i = threadIdx.x; i < kGiniWin; i += kThreads) {
      const unsigned int c = s_hist[i];
      if (c) atomicAdd(ed.hist + win_lo + i, c);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kEcoF; ++i) {
    const float w = warp_sum(fs[i]);
    if (lane == 0) s_red[warp * kEcoRow + i] = (double)w;
  }
  {
    const int w = warp_sum(emp);
    if (lane == 0) s_red[warp * kEcoRow + kEcoF] = (double)w;
    // the range as two maxima: the highest bin, and minus the lowest one
    const double hi = warp_max((double)bin_hi), nlo = warp_max(-(double)bin_lo);
    if (lane == 0) { s_red[warp * kEcoRow + kEcoAcc] = hi; s_red[warp * kEcoRow + kEcoAcc + 1] = nlo; }
  }
  __syncthreads();
  if (threadIdx.x < kEcoRow) {
    const bool is_max = threadIdx.x >= kEcoAcc;
    double r = s_red[threadIdx.x];
    for (int w = 1; w < kThreads / 32; ++w) { const double v = s_red[w * kEcoRow + threadIdx.x]; r = is_max ? fmax(r, v) : r + v; }
    ed.partials[(size_t)(row_offset + blockIdx.x) * kEcoRow + threadIdx.x] = r;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&md.ctrl->ticket, 1u) == (unsigned)total_ctas - 1u);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = warp; i < kEcoRow; i += kThreads / 32) {
    const bool is_max = i >= kEcoAcc;
    double r = is_max ? -1.0 / 0.0 : 0.0;
    for (int b = lane; b < total_ctas; b += 32) { const double v = __ldcg(ed.partials + (size_t)b * kEcoRow + i); r = is_max ? fmax(r, v) : r + v; }
    r = is_max ? warp_max(r) : warp_sum(r);
    if (lane == 0) s_tot[i] = r;
  }
  __syncthreads();
  // sharded population: fold the ranks' rows over NVLink peer memory (every rank then runs the same
  // update_environment on identical totals).  Its flags also order the histograms: every household of every
  // rank has counted its income into its rank's histogram before any rank leaves this exchange.
  if (md.world_size > 1) peer_exchange(md, s_tot, kEcoRow, kEcoAcc, kEcoRow);
  if (threadIdx.x == 0) {
    md.ctrl->ticket = 0;
    const Key uk = {kp[2 * md.n_types], kp[2 * md.n_types + 1]};
    eco_update_environment<MODE>(md, s_tot, md.env, uk);
    // scan tiles that hold an income of this step (the whole population's): all the Gini kernels skip the rest
    const double hi = s_tot[kEcoAcc], lo = -s_tot[kEcoAcc + 1];
    const bool any = hi >= lo;
    ed.tile_range[0] = any ? (int)lo / kGiniScanTile : 1;
    ed.tile_range[1] = any ? (int)hi / kGiniScanTile : 0;
  }
}

// ---------------------------------------------------------------------------------------
// Gini (:1789-1809): 2 * sum(index * sorted incomes) / (n * sum) - (n+1)/n without sorting.
// rank(x_i) = 1 + #{j : key_j in a lower bin} + (mean position inside its own bin); the bins are
// the top 22 bits of the order-preserving integer image of the float, i.e. 2^-13 relative width,
// so incomes that share a bin are interchangeable to ~1e-9 of the rank-weighted sum.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int gini_bin(float x) {
  unsigned int b = __float_as_uint(x);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return b >> (32 - kGiniBits);
}

// sharded population: the counts of the whole population = sum over the ranks' histograms, read straight out of
// the peers' exchange allocations (CUDA IPC / NVLink peer loads; every rank does the whole fold over the populated
// tiles -- a few hundred KB per peer).  One CTA per scan tile.
__global__ void __launch_bounds__(kThreads) gini_gather_kernel(const ModelDev md, const EcoDev ed) {
  const int tile = blockIdx.x;
  if (tile < ed.tile_range[0] || tile > ed.tile_range[1]) return;
  const size_t base = (size_t)tile * kGiniScanTile + (size_t)threadIdx.x * 16;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int p = 0; p < md.world_size; ++p) {
      const uint4 v = __ldcv((const uint4*)(xchg_hist(md.xpeer[p]) + base) + j);     // never from a stale L1 line
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    ((uint4*)(ed.bin_count + base))[j] = acc;
  }
}

// this rank's histogram back to zero over the populated tiles (16 MB memset -> a few hundred KB).  Sharded: runs
// after gini_accumulate_kernel, whose tail exchange every rank only leaves once all ranks have gathered.
__global__ void __launch_bounds__(kThreads) gini_clear_kernel(const EcoDev ed) {
  const int tile = blockIdx.x;
  if (tile < ed.tile_range[0] || tile > ed.tile_range[1]) return;
  uint4* p = (uint4*)(ed.hist + (size_t)tile * kGiniScanTile) + threadIdx.x * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) p[j] = make_uint4(0, 0, 0, 0);
}

// exclusive scan of the bin counts: per-tile sums, scan of the sums, per-tile rescan (populated tiles only)
__global__ void __launch_bounds__(kThreads) gini_scan_sums_kernel(const EcoDev ed, const unsigned int* bin_count, unsigned int* sums) {
  __shared__ unsigned int s_w[kThreads / 32];
  if ((int)blockIdx.x < ed.tile_range[0] || (int)blockIdx.x > ed.tile_range[1]) {
    if (threadIdx.x == 0) sums[blockIdx.x] = 0u;
    return;
  }
  const uint4* p = (const uint4*)(bin_count + (size_t)blockIdx.x * kGiniScanTile) + threadIdx.x * 4;
  unsigned int s = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) { const uint4 v = p[j]; s += v.x + v.y + v.z + v.w; }
  s = (unsigned int)warp_sum((int)s);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int tot = 0;
    for (int w = 0; w < kThreads / 32; ++w) tot += s_w[w];
    sums[blockIdx.x] = tot;
  }
}

__global__ void __launch_bounds__(1024) gini_scan_top_kernel(unsigned int* sums, int n, int* win_tile) {   // n <= 1024
  __shared__ unsigned int s_w[32];
  __shared__ unsigned int s_v[1025];
  __shared__ unsigned long long s_best[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned int v = tid < n ? sums[tid] : 0u;
  // the two adjacent tiles holding the most incomes: next step's shared-memory histogram window
  s_v[tid] = v;
  if (tid == 0) s_v[1024] = 0u;
  __syncthreads();
  {
    unsigned long long key = tid + 1 < n || n == 1 ? (((unsigned long long)(v + s_v[tid + 1])) << 32) | (unsigned int)(0xFFFFu - tid) : 0ull;
    if (tid >= n) key = 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o); key = k2 > key ? k2 : key; }
    if (lane == 0) s_best[warp] = key;
    __syncthreads();
    if (tid == 0 && win_tile) {
      unsigned long long best = 0ull;
      for (int w = 0; w < 32; ++w) best = s_best[w] > best ? s_best[w] : best;
      *win_tile = (best >> 32) ? (int)(0xFFFFu - (unsigned int)(best & 0xFFFFu)) : -1;
    }
  }
  unsigned int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  unsigned int off = 0;
  for (int w = 0; w < warp; ++w) off += s_w[w];
  if (tid < n) sums[tid] = off + inc - v;
}

__global__ void __launch_bounds__(kThreads) gini_scan_apply_kernel(const EcoDev ed, const unsigned int* bin_count, const unsigned int* sums,
                                                                   unsigned int* bin_base) {
  __shared__ unsigned int s_w[kThreads / 32];
  if ((int)blockIdx.x < ed.tile_range[0] || (int)blockIdx.x > ed.tile_range[1]) return;    // no income looks these bins up
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t base = (size_t)blockIdx.x * kGiniScanTile + (size_t)tid * 16;
  unsigned int c[16];
  unsigned int s = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 v = ((const uint4*)(bin_count + base))[j];
    c[4 * j] = v.x; c[4 * j + 1] = v.y; c[4 * j + 2] = v.z; c[4 * j + 3] = v.w;
    s += v.x + v.y + v.z + v.w;
  }
  unsigned int inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  unsigned int off = sums[blockIdx.x];
  for (int w = 0; w < warp; ++w) off += s_w[w];
  unsigned int run = off + inc - s;
#pragma unroll
  for (int j = 0; j < 16; ++j) { const unsigned int cc = c[j]; c[j] = run; run += cc; }
#pragma unroll
  for (int j = 0; j < 4; ++j) ((uint4*)(bin_base + base))[j] = make_uint4(c[4 * j], c[4 * j + 1], c[4 * j + 2], c[4 * j + 3]);
}

// rank-weighted sum + plain sum of the incomes; last CTA: Gini, compute_metrics, history row,
// and the step bookkeeping (model.py:203-213).
__global__ void __launch_bounds__(kThreads) gini_accumulate_kernel(const ModelDev md, const EcoDev ed, int hh_type) {
  __shared__ double s_a[kThreads / 32], s_b[kThreads / 32];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double wsum = 0.0, xsum = 0.0;
  long long n = 0;
  if (hh_type >= 0) {
    const float* income = (const float*)md.t[hh_type].f[1];
    n = md.t[hh_type].n;
    for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < n; i += (long long)gridDim.x * blockDim.x) {
      const float x = __ldg(income + i);
      const unsigned int b = gini_bin(x);
      const unsigned int base = __ldg(ed.bin_base + b), c = __ldg(ed.bin_count + b);
      // index * income is a float32 product in the reference (:1801); the index of a tie group is its mean
      const float idx = (float)((double)base + 0.5 * ((double)c + 1.0));
      wsum += (double)(idx * x);
      xsum += (double)x;
    }
  }
  wsum = warp_sum(wsum); xsum = warp_sum(xsum);
  if (lane == 0) { s_a[warp] = wsum; s_b[warp] = xsum; }
  __syncthreads();
  if (tid == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < kThreads / 32; ++w) { a += s_a[w]; b += s_b[w]; }
    ed.gini_partials[2 * blockIdx.x] = a;
    ed.gini_partials[2 * blockIdx.x + 1] = b;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(ed.ticket2, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  __shared__ double s_gt[2];
  if (warp == 0) {
    double a = 0, b = 0;
    for (int i = lane; i < (int)gridDim.x; i += 32) { a += __ldcg(ed.gini_partials + 2 * i); b += __ldcg(ed.gini_partials + 2 * i + 1); }
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) { s_gt[0] = a; s_gt[1] = b; *ed.ticket2 = 0; }
  }
  __syncthreads();
  // sharded population: ranks hold the GLOBAL histogram (gini_gather_kernel), so their partial rank-weighted
  // sums simply add up
  if (md.world_size > 1) peer_exchange(md, s_gt, 2, 0, 0);
  if (tid == 0) {
    const double a = s_gt[0], b = s_gt[1];
    float gini = 0.0f;
    if (hh_type >= 0 && md.t[hh_type].gn > 0) {
      const float income_sum = (float)b, weighted = (float)a;
      if (income_sum > 1e-6f) {
        const float nf = (float)md.t[hh_type].gn;
        gini = (2.0f * weighted) / (nf * income_sum) - (float)(((double)md.t[hh_type].gn + 1.0) / (double)md.t[hh_type].gn);
      }
    }
    md.env[EE_GINI] = gini;
    Ctrl* c = md.ctrl;
    const long long t = c->time_step + 1;
    if ((t % md.collect_interval) == 0) {
      double m[kMaxMetrics];
#pragma unroll
      for (int i = 0; i < kMaxMetrics; ++i) m[i] = 0.0;
      eco_compute_metrics(md.env, gini, m);
      double* row = md.metrics + (size_t)c->n_recorded * kMaxMetrics;
#pragma unroll
      for (int i = 0; i < kMaxMetrics; ++i) row[i] = m[i];
      md.record_steps[c->n_recorded] = (int)t;
      c->n_recorded += 1;
    }
    c->time_step = t;
    c->step_in_run += 1;
  }
}

}  // namespace jxb
