// prng.cuh -- counter-based threefry2x32 and the JAX key algebra built on it.
//
// Bit-compatible with jax.random (jax/_src/prng.py) in both stream layouts; the
// reference calls it at jaxabm/model.py:46,129,156,164,183 and jaxabm/agent.py:115,156.
// Everything here is O(1) per agent: a child key of split(key, N) is derived in
// registers, never materialised as a 2N-word array in HBM.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define JXB_HD __host__ __device__ __forceinline__
#else
#define JXB_HD inline
#endif

namespace jxb {

struct Key {
  uint32_t a, b;
};

JXB_HD uint32_t rotl32(uint32_t x, int r) {
#ifdef __CUDA_ARCH__
  return __funnelshift_l(x, x, r);
#else
  return (x << r) | (x >> (32 - r));
#endif
}

// Threefry-2x32, 20 rounds (Random123).  ks2 = k0 ^ k1 ^ 0x1BD11BDA.
JXB_HD void threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t& y0,
                         uint32_t& y1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += k0;
  x1 += k1;
#define JXB_TF_R(r)  \
  x0 += x1;          \
  x1 = rotl32(x1, r); \
  x1 ^= x0;
  JXB_TF_R(13) JXB_TF_R(15) JXB_TF_R(26) JXB_TF_R(6)
  x0 += k1; x1 += k2 + 1u;
  JXB_TF_R(17) JXB_TF_R(29) JXB_TF_R(16) JXB_TF_R(24)
  x0 += k2; x1 += k0 + 2u;
  JXB_TF_R(13) JXB_TF_R(15) JXB_TF_R(26) JXB_TF_R(6)
  x0 += k0; x1 += k1 + 3u;
  JXB_TF_R(17) JXB_TF_R(29) JXB_TF_R(16) JXB_TF_R(24)
  x0 += k1; x1 += k2 + 4u;
  JXB_TF_R(13) JXB_TF_R(15) JXB_TF_R(26) JXB_TF_R(6)
  x0 += k2; x1 += k0 + 5u;
#undef JXB_TF_R
  y0 = x0;
  y1 = x1;
}

// word `f` of threefry_2x32(key, iota(m)) in the ORIGINAL layout: the count vector is
// padded to even length 2h, halved into x0=[0..h), x1=[h..2h), and the outputs are
// concatenated (y0 then y1).
JXB_HD uint32_t original_word(Key k, uint64_t f, uint64_t m) {
  const uint64_t h = (m + 1) >> 1;
  const bool odd = (m & 1);
  uint32_t y0, y1;
  if (f < h) {
    uint64_t c1 = h + f;
    uint32_t x1 = (odd && c1 == m) ? 0u : (uint32_t)c1;
    threefry2x32(k.a, k.b, (uint32_t)f, x1, y0, y1);
    return y0;
  }
  uint64_t i = f - h;
  uint32_t x1 = (odd && f == m) ? 0u : (uint32_t)f;
  threefry2x32(k.a, k.b, (uint32_t)i, x1, y0, y1);
  return y1;
}

// child j of jax.random.split(key, n).
template <int MODE>
JXB_HD Key split_child(Key k, uint64_t j, uint64_t n) {
  Key out;
  if (MODE == 1) {  // partitionable: TF(key; hi(j), lo(j)), both words
    threefry2x32(k.a, k.b, (uint32_t)(j >> 32), (uint32_t)j, out.a, out.b);
  } else {  // original: words 2j and 2j+1 of the 2n-word stream
    const uint64_t f0 = 2 * j, f1 = 2 * j + 1;
    if (f1 < n) {  // both in the y0 half, different blocks
      out.a = original_word(k, f0, 2 * n);
      out.b = original_word(k, f1, 2 * n);
    } else if (f0 >= n) {
      out.a = original_word(k, f0, 2 * n);
      out.b = original_word(k, f1, 2 * n);
    } else {  // f0 = n-1 (y0 half), f1 = n (y1 half)
      out.a = original_word(k, f0, 2 * n);
      out.b = original_word(k, f1, 2 * n);
    }
  }
  return out;
}

JXB_HD Key split_child(int mode, Key k, uint64_t j, uint64_t n) {
  return mode == 1 ? split_child<1>(k, j, n) : split_child<0>(k, j, n);
}

// element j of jax.random.bits(key, (m,)) (32-bit).
template <int MODE>
JXB_HD uint32_t bits_elem(Key k, uint64_t j, uint64_t m) {
  if (MODE == 1) {
    uint32_t y0, y1;
    threefry2x32(k.a, k.b, (uint32_t)(j >> 32), (uint32_t)j, y0, y1);
    return y0 ^ y1;
  }
  return original_word(k, j, m);
}

// jax.random.bits(key, ()) -- a scalar draw.
template <int MODE>
JXB_HD uint32_t bits_scalar(Key k) {
  uint32_t y0, y1;
  threefry2x32(k.a, k.b, 0u, 0u, y0, y1);
  return MODE == 1 ? (y0 ^ y1) : y0;
}

JXB_HD float bits_to_unit(uint32_t bits) {
  uint32_t fb = (bits >> 9) | 0x3F800000u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(fb) - 1.0f;
#else
  union { uint32_t u; float f; } c;
  c.u = fb;
  return c.f - 1.0f;
#endif
}

// jax.random.uniform: max(lo, u*(hi-lo)+lo), float32, no FMA contraction (-fmad=false).
JXB_HD float bits_to_uniform(uint32_t bits, float lo, float hi) {
  float u = bits_to_unit(bits);
  float v = u * (hi - lo) + lo;
  return v > lo ? v : lo;
}

// Keyed bijection of [0, n): balanced Feistel over the smallest even bit width covering
// n, 4 rounds, cycle-walking.  Restated in oracle/jaxlike.py::feistel_permute.
JXB_HD uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}

struct Feistel {
  uint32_t n, half, mask;
  uint32_t rk[4];
};

JXB_HD Feistel make_feistel(uint32_t n, const uint32_t* rk) {
  Feistel f;
  f.n = n;
  uint32_t b = 2;
  if (n > 1) {
    uint32_t v = n - 1, bl = 0;
    while (v) { ++bl; v >>= 1; }
    b = bl < 2 ? 2 : bl;
  }
  b += (b & 1);
  f.half = b / 2;
  f.mask = (1u << f.half) - 1u;
  for (int i = 0; i < 4; ++i) f.rk[i] = rk[i];
  return f;
}

// the inverse bijection: rounds undone in reverse order, cycle-walking through the same out-of-range values
JXB_HD uint32_t feistel_inverse(const Feistel& f, uint32_t idx) {
  if (f.n <= 1) return idx;
  uint32_t v = idx;
  do {
    uint32_t l = v >> f.half, r = v & f.mask;
#pragma unroll
    for (int i = 3; i >= 0; --i) {
      uint32_t t = r ^ (mix32(l ^ f.rk[i]) & f.mask);
      r = l;
      l = t;
    }
    v = (l << f.half) | r;
  } while (v >= f.n);
  return v;
}

JXB_HD uint32_t feistel_permute(const Feistel& f, uint32_t idx) {
  if (f.n <= 1) return idx;
  uint32_t v = idx;
  do {
    uint32_t l = v >> f.half, r = v & f.mask;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t t = l ^ (mix32(r ^ f.rk[i]) & f.mask);
      l = r;
      r = t;
    }
    v = (l << f.half) | r;
  } while (v >= f.n);
  return v;
}

// Several walks per thread, advanced in ONE loop: a lane moves on to its next entry as soon as the current one lands
// inside [0, n), so a warp runs for the slowest lane's TOTAL number of applications instead of the sum over the
// entries of the slowest lane of each (cycle walking over a domain up to 4x the range diverges badly otherwise).
// Entries are j0 + i * stride, i < cnt; k[i] = feistel_inverse(f, j0 + i * stride).
template <int N>
__device__ __forceinline__ void feistel_inverse_strided(const Feistel& f, uint32_t j0, uint32_t stride, int cnt, uint32_t (&k)[N]) {
#pragma unroll
  for (int q = 0; q < N; ++q) k[q] = j0 + q * stride;
  if (f.n <= 1) return;
  int i = 0;
  uint32_t v = j0;
  while (i < cnt) {
    uint32_t l = v >> f.half, r = v & f.mask;
#pragma unroll
    for (int q = 3; q >= 0; --q) {
      uint32_t t = r ^ (mix32(l ^ f.rk[q]) & f.mask);
      r = l;
      l = t;
    }
    v = (l << f.half) | r;
    if (v < f.n) {
#pragma unroll
      for (int q = 0; q < N; ++q)
        if (i == q) k[q] = v;
      ++i;
      v = j0 + i * stride;
    }
  }
}

// x[i] <- feistel_permute(f, x[i]) for the entries whose bit is set in `todo`, advanced in one flat loop as above
template <int N>
__device__ __forceinline__ void feistel_permute_masked(const Feistel& f, unsigned int todo, uint32_t (&x)[N]) {
  if (f.n <= 1) return;
  int i = __ffs(todo) - 1;
  uint32_t v = 0;
#pragma unroll
  for (int q = 0; q < N; ++q)
    if (i == q) v = x[q];
  while (todo) {
    uint32_t l = v >> f.half, r = v & f.mask;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t t = l ^ (mix32(r ^ f.rk[q]) & f.mask);
      l = r;
      r = t;
    }
    v = (l << f.half) | r;
    if (v < f.n) {
#pragma unroll
      for (int q = 0; q < N; ++q)
        if (i == q) x[q] = v;
      todo &= todo - 1;
      i = __ffs(todo) - 1;
#pragma unroll
      for (int q = 0; q < N; ++q)
        if (i == q) v = x[q];
    }
  }
}

}  // namespace jxb
