"""Host-side ``jax.random`` key algebra, bit-compatible, served by ``libjxb.so``.

The reference does this scalar work on the host through JAX
(``jaxabm/model.py:46,129,156``, ``jaxabm/analysis.py:61-95,438-441``,
``jaxabm/agentpy.py:509-513``).  Keys are ``uint32[2]`` NumPy arrays.  Per-agent key
derivation never happens here -- it is done in registers inside the CUDA kernels.
"""
from __future__ import annotations

import numpy as np

from . import _native as nat

__all__ = ["PRNGKey", "split", "bits", "uniform", "normal", "randint", "permutation", "threefry2x32"]


def _traced(key) -> bool:
    from .trace import TrKey
    return isinstance(key, TrKey)


def _mode(mode):
    return nat.default_rng_mode() if mode is None else mode


def _key(key) -> np.ndarray:
    k = np.ascontiguousarray(np.asarray(key, dtype=np.uint32).reshape(2))
    return k


def PRNGKey(seed: int) -> np.ndarray:
    """``jax.random.PRNGKey(seed)`` for a 32-bit seed: ``[0, uint32(seed)]``."""
    seed = int(seed)
    return np.array([0xFFFFFFFF if seed < 0 else 0, seed & 0xFFFFFFFF], dtype=np.uint32)


def threefry2x32(key, ctr) -> np.ndarray:
    out = np.zeros(2, dtype=np.uint32)
    c = np.ascontiguousarray(np.asarray(ctr, dtype=np.uint32).reshape(2))
    nat.check(nat.lib().jxb_prng_threefry2x32(nat.ptr(_key(key)), nat.ptr(c), nat.ptr(out)))
    return out


def split(key, num: int = 2, mode=None) -> np.ndarray:
    if _traced(key):                      # inside a traced rule: symbolic children (jaxabm_b200/trace.py)
        from .trace import key_split
        return key_split(key, num)
    out = np.zeros((num, 2), dtype=np.uint32)
    nat.check(nat.lib().jxb_prng_split(_mode(mode), nat.ptr(_key(key)), int(num), nat.ptr(out)))
    return out


def bits(key, shape=(), mode=None) -> np.ndarray:
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    n = int(np.prod(shape)) if shape else 1
    out = np.zeros(n, dtype=np.uint32)
    nat.check(nat.lib().jxb_prng_bits(_mode(mode), nat.ptr(_key(key)), n, nat.ptr(out)))
    return out.reshape(shape)


def normal(key, shape=(), mode=None):
    """``jax.random.normal`` -- traced rules only (host code of this package never draws normals)."""
    if _traced(key):
        from .trace import key_normal
        return key_normal(key, shape)
    raise NotImplementedError("random.normal is available inside traced rules only")


def uniform(key, shape=(), minval=0.0, maxval=1.0, mode=None) -> np.ndarray:
    if _traced(key):
        from .trace import key_uniform
        return key_uniform(key, shape, minval, maxval)
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    n = int(np.prod(shape)) if shape else 1
    out = np.zeros(n, dtype=np.float32)
    nat.check(nat.lib().jxb_prng_uniform(_mode(mode), nat.ptr(_key(key)), n, float(minval), float(maxval),
                                         nat.ptr(out)))
    return out.reshape(shape)


def randint(key, shape, minval: int, maxval: int, mode=None) -> np.ndarray:
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    n = int(np.prod(shape)) if shape else 1
    out = np.zeros(n, dtype=np.int32)
    nat.check(nat.lib().jxb_prng_randint(_mode(mode), nat.ptr(_key(key)), n, int(minval), int(maxval),
                                         nat.ptr(out)))
    return out.reshape(shape)


def permutation(key, x, mode=None) -> np.ndarray:
    """``jax.random.permutation`` of a 1-D array: rounds of stable sort by fresh random bits."""
    x = np.asarray(x).copy()
    n = x.size
    rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))
    for _ in range(rounds):
        key, sub = split(key, 2, mode)
        x = x[np.argsort(bits(sub, (n,), mode), kind="stable")]
    return x
