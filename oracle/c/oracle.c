/*
 * oracle.c -- C/OpenMP restatement of the builder-authored Schelling step and of the
 * market / walker updates, for (a) the multi-core CPU baseline of bench.py and (b)
 * full-size parity checks that the NumPy oracle is too slow for.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Follows: oracle/rules.py::SchellingAgent.update_collective + schelling_update_state
 * (rule of DESIGN.md, layout of examples/models/schelling_model.py:26-31,119-139);
 * jax.random semantics as in oracle/jaxlike.py (jax/_src/prng.py, both stream layouts);
 * tests/integration/test_integration.py:43-67,94-121,125-160 for the market step.
 * It is validated against the NumPy oracle in tests/test_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static void threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* y0, uint32_t* y1) {
  static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0];
  x1 += ks[1];
  for (int i = 0; i < 5; ++i) {
    for (int j = 0; j < 4; ++j) {
      x0 += x1;
      x1 = rotl32(x1, R[i % 2][j]);
      x1 ^= x0;
    }
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
  }
  *y0 = x0;
  *y1 = x1;
}

/* element j of jax.random.bits(key, (m,)) */
static uint32_t bits_elem(int mode, const uint32_t key[2], uint64_t j, uint64_t m) {
  uint32_t y0, y1;
  if (mode == 1) {
    threefry2x32(key[0], key[1], (uint32_t)(j >> 32), (uint32_t)j, &y0, &y1);
    return y0 ^ y1;
  }
  uint64_t h = (m + 1) / 2;
  int odd = (int)(m & 1);
  if (j < h) {
    uint64_t c1 = h + j;
    threefry2x32(key[0], key[1], (uint32_t)j, (odd && c1 == m) ? 0u : (uint32_t)c1, &y0, &y1);
    return y0;
  }
  threefry2x32(key[0], key[1], (uint32_t)(j - h), (uint32_t)j, &y0, &y1);
  return y1;
}

static void split_child(int mode, const uint32_t key[2], uint64_t j, uint64_t n, uint32_t out[2]) {
  if (mode == 1) {
    threefry2x32(key[0], key[1], (uint32_t)(j >> 32), (uint32_t)j, &out[0], &out[1]);
  } else {
    out[0] = bits_elem(0, key, 2 * j, 2 * n);
    out[1] = bits_elem(0, key, 2 * j + 1, 2 * n);
  }
}

static inline float bits_to_unit(uint32_t b) {
  union { uint32_t u; float f; } c;
  c.u = (b >> 9) | 0x3F800000u;
  return c.f - 1.0f;
}

static inline uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}

typedef struct { uint32_t n, half, mask, rk[4]; } feistel_t;

static feistel_t make_feistel(uint32_t n, const uint32_t* rk) {
  feistel_t f;
  f.n = n;
  uint32_t b = 2;
  if (n > 1) { uint32_t v = n - 1, bl = 0; while (v) { ++bl; v >>= 1; } b = bl < 2 ? 2 : bl; }
  b += (b & 1);
  f.half = b / 2;
  f.mask = (1u << f.half) - 1u;
  for (int i = 0; i < 4; ++i) f.rk[i] = rk[i];
  return f;
}

static uint32_t feistel_permute(const feistel_t* f, uint32_t idx) {
  if (f->n <= 1) return idx;
  uint32_t v = idx;
  do {
    uint32_t l = v >> f->half, r = v & f->mask;
    for (int i = 0; i < 4; ++i) { uint32_t t = l ^ (mix32(r ^ f->rk[i]) & f->mask); l = r; r = t; }
    v = (l << f->half) | r;
  } while (v >= f->n);
  return v;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline asks for the host's cores explicitly */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

void orc_threefry2x32(const uint32_t key[2], const uint32_t ctr[2], uint32_t out[2]) {
  threefry2x32(key[0], key[1], ctr[0], ctr[1], &out[0], &out[1]);
}

void orc_split(int mode, const uint32_t key[2], int n, uint32_t* out) {
  for (int j = 0; j < n; ++j) split_child(mode, key, (uint64_t)j, (uint64_t)n, out + 2 * j);
}

/* The scalar key chain of jaxabm/model.py:129-130,156,164,183 for `steps` steps:
 * rng in/out; coll_keys [steps][C][2]; update_keys [steps][2].                     */
void orc_key_schedule(int mode, uint32_t rng[2], int C, int has_env_fn, int steps, uint32_t* coll_keys,
                      uint32_t* update_keys) {
  for (int t = 0; t < steps; ++t) {
    uint32_t k[2], sk[2];
    split_child(mode, rng, 1, 2, sk);
    split_child(mode, rng, 0, 2, k);
    rng[0] = k[0]; rng[1] = k[1];
    for (int c = 0; c < C; ++c) {
      uint32_t ck[2], nk[2];
      split_child(mode, sk, 1, 2, ck);
      split_child(mode, sk, 0, 2, nk);
      sk[0] = nk[0]; sk[1] = nk[1];
      coll_keys[((size_t)t * C + c) * 2] = ck[0];
      coll_keys[((size_t)t * C + c) * 2 + 1] = ck[1];
    }
    uint32_t uk[2] = {0, 0};
    if (has_env_fn) split_child(mode, sk, 1, 2, uk);
    update_keys[2 * t] = uk[0];
    update_keys[2 * t + 1] = uk[1];
  }
}

/* ------------------------------------------------------------------------------------------
 * One Schelling step (DESIGN.md "Schelling rule").
 *   grid  int32[W*H]  in: pre-step grid (-1 empty / type), out: grid rebuilt from the moved agents
 *   type  int32[n], pos int32[n][2] (in/out), satisfied uint8[n] (out), moves int32[n] (in/out)
 *   E     int32[e]    in/out: env['empty_cells'] as cell ids in slot order (a mover's old cell
 *                     takes the slot of the cell it moved to)
 *   scratch: cell_agent int32[W*H], U int32[n]   (caller-allocated)
 * Returns the number of movers; *seg_sum / *seg_cnt give sum and count of same/occupied over
 * agents with at least one occupied neighbour (pre-step grid); *n_unsat the unsatisfied count.
 * ------------------------------------------------------------------------------------------ */
int64_t orc_schelling_step(int W, int H, int periodic, float thr, int mode, const uint32_t coll_key[2],
                           int64_t n, const int32_t* type, int32_t* pos, uint8_t* satisfied, int32_t* moves,
                           int32_t* grid, int32_t* cell_agent, int32_t* U, int32_t* E, int64_t e,
                           double* seg_sum, int64_t* seg_cnt, int64_t* n_unsat) {
  const int64_t cells = (int64_t)W * H;
  double ssum = 0.0;
  int64_t scnt = 0;
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < cells; ++c) cell_agent[c] = -1;
#pragma omp parallel for schedule(static) reduction(+ : ssum, scnt)
  for (int64_t i = 0; i < n; ++i) {
    const int x = pos[2 * i], y = pos[2 * i + 1];
    int occ = 0, same = 0;
    for (int dx = -1; dx <= 1; ++dx)
      for (int dy = -1; dy <= 1; ++dy) {
        if (!dx && !dy) continue;
        int xx = x + dx, yy = y + dy;
        if (periodic) {
          xx = (xx + W) % W;
          yy = (yy + H) % H;
        } else if (xx < 0 || xx >= W || yy < 0 || yy >= H) {
          continue;
        }
        const int v = grid[(int64_t)xx * H + yy];
        if (v >= 0) { ++occ; same += (v == type[i]); }
      }
    float frac = 0.f;
    if (occ) { frac = (float)same / (float)occ; ssum += (double)frac; ++scnt; }
    satisfied[i] = (uint8_t)(occ == 0 || frac >= thr);
    cell_agent[(int64_t)x * H + y] = (int32_t)i;
  }
  /* ordered compaction: unsatisfied agents by ascending cell id */
  int64_t u = 0;
  for (int64_t c = 0; c < cells; ++c) {
    const int a = cell_agent[c];
    if (a >= 0 && !satisfied[a]) U[u++] = a;
  }
  const int64_t m = u < e ? u : e;
  uint32_t rk[8];
  for (int i = 0; i < 8; ++i) rk[i] = bits_elem(mode, coll_key, (uint64_t)i, 8);
  const feistel_t fu = make_feistel((uint32_t)u, rk), fe = make_feistel((uint32_t)e, rk + 4);
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < m; ++k) {
    const int a = U[feistel_permute(&fu, (uint32_t)k)];
    const uint32_t j = feistel_permute(&fe, (uint32_t)k);
    const int64_t dst = E[j];
    const int64_t src = (int64_t)pos[2 * a] * H + pos[2 * a + 1];
    E[j] = (int32_t)src;
    grid[dst] = type[a];
    grid[src] = -1;
    pos[2 * a] = (int32_t)(dst / H);
    pos[2 * a + 1] = (int32_t)(dst % H);
    moves[a] += 1;
  }
  *seg_sum = ssum;
  *seg_cnt = scnt;
  *n_unsat = u;
  return m;
}

/* ------------------------------------------------------------------------------------------
 * One consumer/producer market step (tests/integration/test_integration.py:43-67,94-121,
 * 125-160), float32 in the reference's operation order; sums are pairwise per thread chunk.
 * env: price_level, gdp, unemployment, total_consumption, total_production (in/out).
 * ------------------------------------------------------------------------------------------ */
void orc_market_step(int64_t nc, float* savings, float* consumption, float* utility, const float* income,
                     float ptc, int64_t np_, float* capital, float* production, float* profit, float prd,
                     float rr, float rate, float* env, double* sum_utility, double* sum_profit) {
  const float price = env[0];
  double tc = 0, tp = 0, su = 0, sp = 0;
#pragma omp parallel for schedule(static) reduction(+ : tc, su)
  for (int64_t i = 0; i < nc; ++i) {
    const float c = ptc * income[i] / price;
    savings[i] = savings[i] + (income[i] - c * price);
    consumption[i] = c;
    utility[i] = logf(c + 1.0f);
    tc += c;
    su += utility[i];
  }
#pragma omp parallel for schedule(static) reduction(+ : tp, sp)
  for (int64_t i = 0; i < np_; ++i) {
    const float q = prd * powf(capital[i], 0.7f);
    const float revenue = q * price;
    const float costs = 0.1f * capital[i] + 0.05f * q;
    const float pf = revenue - costs;
    capital[i] = capital[i] + rr * pf;
    production[i] = q;
    profit[i] = pf;
    tp += q;
    sp += pf;
  }
  const float total_c = (float)tc, total_p = (float)tp;
  const float ratio = (total_p + 1e-8f) / (total_c + 1e-8f);
  const float change = rate * (1.0f - ratio);
  float p = price * (1.0f + change);
  p = fminf(fmaxf(p, 0.5f), 2.0f);
  env[0] = p;
  env[1] = total_p * p;
  env[2] = fmaxf(0.0f, fminf(0.5f, 1.0f - ratio));
  env[3] = total_c;
  env[4] = total_p;
  *sum_utility = su;
  *sum_profit = sp;
}

/* ------------------------------------------------------------------------------------------
 * One SIR step (DESIGN.md "SIR rule"; oracle/rules.py::SIRAgent.update_batch): synchronous on the
 * pre-step snapshot `st`, k = infected neighbours in the CSR, u = uniform(split(coll_key, N)[i])
 * (jaxabm/agent.py:156), S -> I iff u < 1 - q[min(k, 4095)], I -> R iff u < gamma.
 * counts[3] = S / I / R after the step.
 * ------------------------------------------------------------------------------------------ */
void orc_sir_step(int mode, const uint32_t coll_key[2], int64_t n, const int64_t* row_ptr, const int32_t* col,
                  const int32_t* st, int32_t* out, const float* q, float gamma, int64_t counts[3]) {
  int64_t cS = 0, cI = 0, cR = 0;
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : cS, cI, cR)
  for (int64_t i = 0; i < n; ++i) {
    int s = st[i];
    int64_t k = 0;
    if (s == 0)
      for (int64_t e = row_ptr[i]; e < row_ptr[i + 1]; ++e) k += (st[col[e]] == 1);
    if ((s == 0 && k > 0) || s == 1) {
      uint32_t ak[2];
      split_child(mode, coll_key, (uint64_t)i, (uint64_t)n, ak);
      const float u = bits_to_unit(bits_elem(mode, ak, 0, 1));
      if (s == 0) {
        const float p = 1.0f - q[k < 4095 ? k : 4095];
        if (u < p) s = 1;
      } else if (u < gamma) {
        s = 2;
      }
    }
    out[i] = s;
    cS += (s == 0); cI += (s == 1); cR += (s == 2);
  }
  counts[0] = cS; counts[1] = cI; counts[2] = cR;
}

/* ------------------------------------------------------------------------------------------
 * One random-walker step (examples/basic_example.py:33-69; oracle/rules.py::RandomWalker.step_batch)
 * + the distance metrics of :141-182 on the NEW positions: sum and max of ||pos - 0.5||_2 (float32).
 * ------------------------------------------------------------------------------------------ */
void orc_walk_step(int64_t n, float* pos, float* vel, int32_t* color, int32_t* steps_taken, float lo, float hi,
                   double* sum_dist, float* max_dist) {
  double sd = 0.0;
  float md = 0.0f;
#pragma omp parallel for schedule(static) reduction(+ : sd) reduction(max : md)
  for (int64_t i = 0; i < n; ++i) {
    int any = 0;
    float c[2];
    for (int a = 0; a < 2; ++a) {
      float p = pos[2 * i + a] + vel[2 * i + a];
      const int b = (p <= lo) || (p >= hi);
      vel[2 * i + a] = vel[2 * i + a] * (float)(1 - 2 * b);
      p = fminf(fmaxf(p, lo), hi);
      pos[2 * i + a] = p;
      any |= b;
      c[a] = p - 0.5f;
    }
    if (any) color[i] = 1 - color[i];
    steps_taken[i] += 1;
    const float d = sqrtf(c[0] * c[0] + c[1] * c[1]);
    sd += (double)d;
    md = d > md ? d : md;
  }
  *sum_dist = sd;
  *max_dist = md;
}
