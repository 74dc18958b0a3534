"""NumPy restatement of the ``jax.random`` / ``jax.vmap`` semantics on the hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Third-party module restated: ``jax`` (``jax/_src/prng.py``, ``jax/_src/random.py``),
pinned by the reference only as ``jax>=0.4.1`` (``requirements.txt:8-9``) and
absent from ``/root/reference``.  Reference call sites this serves:

* ``jax.random.PRNGKey``   -- ``jaxabm/model.py:46``, ``jaxabm/analysis.py:61``
* ``jax.random.split``     -- ``jaxabm/model.py:129,156,164,183``; ``jaxabm/agent.py:115,156``
* ``jax.random.uniform``   -- ``tests/integration/test_integration.py:35,87``; ``jaxabm/analysis.py:81``
* ``jax.random.randint``   -- ``jaxabm/agentpy.py:510-512``; ``jaxabm/analysis.py:441``
* ``jax.random.normal``    -- ``examples/models/advanced_economic_model.py:409``
* ``jax.random.permutation`` -- ``jaxabm/analysis.py:86``

Two stream layouts exist in JAX (flag ``jax_threefry_partitionable``; default
False before JAX 0.5.0, True from 0.5.0).  Both are implemented; ``MODE`` picks
the module-wide default and every function takes ``mode=`` explicitly.

Pinned by ``tests/golden/threefry_kat.json`` (Random123 KATs and the published
JAX values for ``split(PRNGKey(0))`` / ``uniform(PRNGKey(0))``).
"""
from __future__ import annotations

import math

import numpy as np

LEGACY = 0
PARTITIONABLE = 1
MODE = PARTITIONABLE

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds (Random123; ``jax/_src/prng.py::_threefry2x32_lowering``).

    All arguments broadcast; returns two uint32 arrays.
    """
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=_U32)
        k1 = np.asarray(k1, dtype=_U32)
        x0 = np.asarray(x0, dtype=_U32).copy()
        x1 = np.asarray(x1, dtype=_U32).copy()
        ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + _U32(i + 1)
    return x0, x1


def PRNGKey(seed: int) -> np.ndarray:
    """``jax.random.PRNGKey`` for a seed that fits 32 bits: ``[0, uint32(seed)]``.

    (x64 disabled: the seed is converted to int32 first, so the high word is 0
    for non-negative seeds and 0xFFFFFFFF for negative ones.)
    """
    seed = int(seed)
    lo = seed & 0xFFFFFFFF
    hi = 0xFFFFFFFF if seed < 0 else 0
    return np.array([hi, lo], dtype=_U32)


def _bits_original(key, m):
    """``threefry_2x32(key, iota(m))``: pad to even, halve, concat, trim."""
    if m == 0:
        return np.zeros(0, dtype=_U32)
    odd = m % 2
    n = (m + odd) // 2
    c = np.arange(m + odd, dtype=np.uint64).astype(_U32)
    if odd:
        c[-1] = 0
    y0, y1 = threefry2x32(key[0], key[1], c[:n], c[n:])
    flat = np.concatenate([y0, y1])
    return flat[:m]


def split(key, num: int = 2, mode: int | None = None) -> np.ndarray:
    """``jax.random.split(key, num)`` -> uint32[num, 2]."""
    mode = MODE if mode is None else mode
    key = np.asarray(key, dtype=_U32)
    if mode == LEGACY:
        return _bits_original(key, 2 * num).reshape(num, 2)
    j = np.arange(num, dtype=np.uint64)
    y0, y1 = threefry2x32(key[0], key[1], (j >> np.uint64(32)).astype(_U32), j.astype(_U32))
    return np.stack([y0, y1], axis=1)


def split_batched(keys, num: int, mode: int | None = None) -> np.ndarray:
    """vmapped ``split``: keys uint32[B,2] -> uint32[B,num,2]."""
    mode = MODE if mode is None else mode
    keys = np.asarray(keys, dtype=_U32)
    B = keys.shape[0]
    k0 = keys[:, 0:1]
    k1 = keys[:, 1:2]
    if mode == LEGACY:
        c = np.arange(2 * num, dtype=_U32)
        y0, y1 = threefry2x32(k0, k1, c[None, :num], c[None, num:])
        return np.concatenate([y0, y1], axis=1).reshape(B, num, 2)
    j = np.arange(num, dtype=_U32)[None, :]
    y0, y1 = threefry2x32(k0, k1, np.zeros_like(j), j)
    return np.stack([y0, y1], axis=2)


def random_bits(key, shape=(), mode: int | None = None) -> np.ndarray:
    """32-bit ``jax.random.bits(key, shape)``."""
    mode = MODE if mode is None else mode
    key = np.asarray(key, dtype=_U32)
    shape = tuple(shape) if not isinstance(shape, int) else (shape,)
    m = int(np.prod(shape)) if shape else 1
    if mode == LEGACY:
        return _bits_original(key, m).reshape(shape)
    j = np.arange(m, dtype=np.uint64)
    y0, y1 = threefry2x32(key[0], key[1], (j >> np.uint64(32)).astype(_U32), j.astype(_U32))
    return (y0 ^ y1).reshape(shape)


def random_bits_scalar_batched(keys, mode: int | None = None) -> np.ndarray:
    """vmapped scalar draw: keys uint32[B,2] -> uint32[B] (``bits(key, ())`` per key)."""
    mode = MODE if mode is None else mode
    keys = np.asarray(keys, dtype=_U32)
    z = np.zeros(keys.shape[0], dtype=_U32)
    y0, y1 = threefry2x32(keys[:, 0], keys[:, 1], z, z)
    return y0 if mode == LEGACY else (y0 ^ y1)


def bits_to_uniform(bits, minval=0.0, maxval=1.0) -> np.ndarray:
    """float32 ``uniform`` from 32 random bits (``jax/_src/random.py::_uniform``)."""
    bits = np.asarray(bits, dtype=_U32)
    fb = (bits >> _U32(9)) | _U32(0x3F800000)
    u = fb.view(np.float32) - np.float32(1.0)
    lo = np.float32(minval)
    hi = np.float32(maxval)
    return np.maximum(lo, u * (hi - lo) + lo).astype(np.float32)


def uniform(key, shape=(), minval=0.0, maxval=1.0, mode: int | None = None) -> np.ndarray:
    return bits_to_uniform(random_bits(key, shape, mode), minval, maxval)


def uniform_scalar_batched(keys, minval=0.0, maxval=1.0, mode: int | None = None) -> np.ndarray:
    return bits_to_uniform(random_bits_scalar_batched(keys, mode), minval, maxval)


def erfinv_f32(x) -> np.ndarray:
    """XLA's float32 ``erf_inv`` (Giles' single-precision polynomial)."""
    x = np.asarray(x, dtype=np.float32)
    f = np.float32
    with np.errstate(divide="ignore", invalid="ignore"):
        w = -np.log1p(-(x * x)).astype(np.float32)
        lt = w < f(5.0)
        wa = (w - f(2.5)).astype(np.float32)
        wb = (np.sqrt(w).astype(np.float32) - f(3.0)).astype(np.float32)
        ca = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087,
              -0.00125372503, -0.00417768164, 0.246640727, 1.50140941)
        cb = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773,
              -0.0076224613, 0.00943887047, 1.00167406, 2.83297682)
        pa = np.full_like(x, f(ca[0]))
        pb = np.full_like(x, f(cb[0]))
        for c in ca[1:]:
            pa = (f(c) + pa * wa).astype(np.float32)
        for c in cb[1:]:
            pb = (f(c) + pb * wb).astype(np.float32)
        p = np.where(lt, pa, pb)
        out = (p * x).astype(np.float32)
        out = np.where(np.abs(x) == f(1.0), np.copysign(f(np.inf), x), out)
    return out.astype(np.float32)


_NORMAL_LO = np.nextafter(np.float32(-1.0), np.float32(0.0))


def bits_to_normal(bits) -> np.ndarray:
    """``jax.random.normal``: sqrt(2) * erfinv(uniform(nextafter(-1,0), 1))."""
    u = bits_to_uniform(bits, _NORMAL_LO, 1.0)
    return (np.float32(np.sqrt(2.0)) * erfinv_f32(u)).astype(np.float32)


def normal(key, shape=(), mode: int | None = None) -> np.ndarray:
    return bits_to_normal(random_bits(key, shape, mode))


def randint(key, shape, minval: int, maxval: int, mode: int | None = None) -> np.ndarray:
    """int32 ``jax.random.randint`` (``jax/_src/random.py::_randint``)."""
    k = split(key, 2, mode)
    hi_bits = random_bits(k[0], shape, mode).astype(_U32)
    lo_bits = random_bits(k[1], shape, mode).astype(_U32)
    span = _U32(1) if maxval <= minval else _U32((int(maxval) - int(minval)) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        mult = _U32((1 << 16) % int(span))
        mult = _U32((int(mult) * int(mult)) & 0xFFFFFFFF) % span
        off = ((hi_bits % span) * mult + (lo_bits % span)) % span
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int32)


def permutation(key, x, mode: int | None = None) -> np.ndarray:
    """``jax.random.permutation(key, x)`` for a 1-D array (sort-based shuffle)."""
    x = np.asarray(x).copy()
    n = x.size
    rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))
    for _ in range(rounds):
        key, sub = split(key, 2, mode)
        sk = random_bits(sub, (n,), mode)
        x = x[np.argsort(sk, kind="stable")]
    return x


# ---------------------------------------------------------------------------
# Keyed bijection on [0, n): used by the builder-authored Schelling mover rule
# (DESIGN.md "Schelling rule").  Not a JAX primitive; restated identically in
# jaxabm_b200/csrc/feistel.cuh.  Balanced Feistel network over the smallest
# even bit width >= log2(n), 4 rounds, cycle-walking back into [0, n).
# ---------------------------------------------------------------------------

def _mix32(x):
    with np.errstate(over="ignore"):
        x = np.asarray(x, dtype=_U32)
        x = x ^ (x >> _U32(16))
        x = x * _U32(0x7FEB352D)
        x = x ^ (x >> _U32(15))
        x = x * _U32(0x846CA68B)
        x = x ^ (x >> _U32(16))
    return x


def feistel_bits(n: int) -> int:
    b = max(2, int(n - 1).bit_length()) if n > 1 else 2
    return b + (b & 1)


def feistel_permute(idx, n: int, rk) -> np.ndarray:
    """pi(idx) for a keyed bijection pi of [0, n); ``rk`` = 4 uint32 round keys."""
    idx = np.asarray(idx, dtype=_U32).copy()
    if n <= 1:
        return idx
    half = feistel_bits(n) // 2
    mask = _U32((1 << half) - 1)
    rk = np.asarray(rk, dtype=_U32)

    def once(v):
        l = v >> _U32(half)
        r = v & mask
        for i in range(4):
            l, r = r, l ^ (_mix32(r ^ rk[i]) & mask)
        return (l << _U32(half)) | r

    out = once(idx)
    bad = out >= _U32(n)
    while bad.any():
        out[bad] = once(out[bad])
        bad = out >= _U32(n)
    return out


# ---------------------------------------------------------------------------
# vmap output broadcasting (jaxabm/agent.py:125-130): outputs that do not depend
# on the mapped key are broadcast to the batch; Python scalars take JAX's
# x64-disabled default dtypes.
# ---------------------------------------------------------------------------

def as_jax_default(value):
    """dtype a Python/NumPy value gets under JAX with x64 disabled."""
    if isinstance(value, (bool, np.bool_)):
        return np.asarray(value, dtype=np.bool_)
    if isinstance(value, (int, np.integer)):
        return np.asarray(value, dtype=np.int32)
    if isinstance(value, float):
        return np.asarray(value, dtype=np.float32)
    a = np.asarray(value)
    if a.dtype == np.float64:
        a = a.astype(np.float32)
    elif a.dtype == np.int64:
        a = a.astype(np.int32)
    return a


def broadcast_to_batch(value, n: int) -> np.ndarray:
    a = as_jax_default(value)
    if a.ndim >= 1 and a.shape[0] == n and getattr(value, "_batched", False):
        return a
    return np.broadcast_to(a, (n,) + a.shape).copy()
