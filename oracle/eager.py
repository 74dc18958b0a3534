"""NumPy batch backend for user-written rules -- TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The traced path of the product (``jaxabm_b200/trace.py``) turns plain user Python into a CUDA kernel.
To check it, the SAME user code is executed here eagerly on whole columns: ``jnp`` / ``random`` below
implement the slice of ``jax.numpy`` / ``jax.random`` that rules use, with JAX's x64-disabled dtype
rules restated independently of the tracer (Python scalars are weak: scalar (op) scalar is Python
double arithmetic, scalar (op) array adopts the array's kind, ``int32 * python float -> float32``;
``bool < int32 < float32``), float32 arithmetic op by op, reductions accumulated in float64 and
rounded once.  ``wrap_agent_type`` / ``wrap_model_fn`` adapt user classes / functions to the batch
protocol of ``oracle.runtime`` (which follows ``jaxabm/agent.py:92-177`` and ``jaxabm/model.py:118-262``).
"""
from __future__ import annotations

import numpy as np

from . import jaxlike as jl

f32, i32 = np.float32, np.int32


class EArr:
    """An eager value: a NumPy array (a column ``(n,)`` or a scalar ``()``) of dtype float32 / int32 / bool."""
    __array_priority__ = 1000

    def __init__(self, v):
        v = np.asarray(v)
        if v.dtype == np.float64:
            v = v.astype(f32)
        elif v.dtype == np.int64:
            v = v.astype(i32)
        elif v.dtype not in (np.dtype(f32), np.dtype(i32), np.dtype(bool)):
            v = v.astype(f32 if v.dtype.kind == "f" else (bool if v.dtype.kind == "b" else i32))
        self.v = v

    @property
    def dtype(self):
        return self.v.dtype

    @property
    def shape(self):
        return self.v.shape

    def astype(self, dt):
        name = getattr(dt, "__name__", str(dt))
        if dt in (float, f32, np.float64) or "float" in name:
            return EArr(self.v.astype(f32))
        if dt in (int, i32, np.int64) or name.startswith("int"):
            return EArr(self.v.astype(i32))
        return EArr(self.v.astype(bool))

    def __bool__(self):
        raise TypeError("truth value of a batched array (use jnp.where)")

    def __add__(self, o): return _bin(np.add, self, o)
    def __radd__(self, o): return _bin(np.add, o, self)
    def __sub__(self, o): return _bin(np.subtract, self, o)
    def __rsub__(self, o): return _bin(np.subtract, o, self)
    def __mul__(self, o): return _bin(np.multiply, self, o)
    def __rmul__(self, o): return _bin(np.multiply, o, self)
    def __truediv__(self, o): return _bin(np.divide, self, o, true_div=True)
    def __rtruediv__(self, o): return _bin(np.divide, o, self, true_div=True)
    def __pow__(self, o): return _pow(self, o)
    def __rpow__(self, o): return _pow(o, self)
    def __neg__(self): return EArr(-_arith(self.v))
    def __abs__(self): return EArr(np.abs(_arith(self.v)))
    def __lt__(self, o): return _cmp(np.less, self, o)
    def __le__(self, o): return _cmp(np.less_equal, self, o)
    def __gt__(self, o): return _cmp(np.greater, self, o)
    def __ge__(self, o): return _cmp(np.greater_equal, self, o)
    def __eq__(self, o): return _cmp(np.equal, self, o)          # noqa: E711
    def __ne__(self, o): return _cmp(np.not_equal, self, o)
    __hash__ = object.__hash__
    def __and__(self, o): return EArr(np.logical_and(self.v, _val(o)))
    __rand__ = __and__
    def __or__(self, o): return EArr(np.logical_or(self.v, _val(o)))
    __ror__ = __or__
    def __invert__(self): return EArr(np.logical_not(self.v))


def _val(x):
    return x.v if isinstance(x, EArr) else x


def _arith(v):
    return v.astype(i32) if v.dtype == bool else v


def _is_py(x):
    return isinstance(x, (bool, int, float)) and not isinstance(x, (np.generic,))


def _target(a, b, true_div=False):
    """Result dtype of a binary arithmetic op under JAX's weak-type rules (None: both Python scalars)."""
    pa, pb = _is_py(a), _is_py(b)
    if pa and pb:
        return None
    def kind(x):
        if _is_py(x):
            return "f" if isinstance(x, float) else ("b" if isinstance(x, bool) else "i")
        return np.asarray(_val(x)).dtype.kind
    ka, kb = kind(a), kind(b)
    if true_div or "f" in (ka, kb):
        return f32
    if pa or pb:                                      # python int / bool with an int / bool array
        return i32
    return i32 if "i" in (ka, kb) else bool


def _bin(fn, a, b, true_div=False):
    dt = _target(a, b, true_div)
    if dt is None:
        return fn(a, b).item() if isinstance(fn(a, b), np.generic) else fn(a, b)
    if dt == bool:
        dt = i32
    x = np.asarray(_val(a)).astype(dt)
    y = np.asarray(_val(b)).astype(dt)
    with np.errstate(all="ignore"):
        return EArr(fn(x, y).astype(dt))


def _cmp(fn, a, b):
    dt = _target(a, b) or f32
    if dt == bool:
        dt = i32
    return EArr(fn(np.asarray(_val(a)).astype(dt), np.asarray(_val(b)).astype(dt)))


def _pow(a, b):
    if _is_py(b) and isinstance(b, int) and not isinstance(b, bool) and 0 <= b <= 64:
        # lax.integer_pow (jax/_src/lax/lax.py _integer_pow_jvp / the XLA lowering): square-and-multiply,
        # x**4 = (x*x)*(x*x); x**0 = ones_like(x), also for inf / nan
        if b == 0:
            if _is_py(a):
                return 1.0 if isinstance(a, float) else 1
            v = np.asarray(_val(a))
            return EArr(np.ones_like(v))
        acc, base, n = None, a, b
        while n > 0:
            if n & 1:
                acc = base if acc is None else acc * base
            n >>= 1
            if n > 0:
                base = base * base
        return acc
    if _is_py(a) and _is_py(b):
        return float(a) ** float(b)
    with np.errstate(all="ignore"):
        return EArr(np.power(np.asarray(_val(a)).astype(f32), np.asarray(_val(b)).astype(f32)).astype(f32))


def _fsum(v):
    return f32(np.sum(np.asarray(v, dtype=np.float64)))


class _JNP:
    float32, int32, bool_ = f32, i32, np.bool_

    def where(self, c, a, b):
        dt = _target(a, b)
        if dt is None:
            dt = f32 if isinstance(a, float) or isinstance(b, float) else (bool if isinstance(a, bool) and isinstance(b, bool) else i32)
        return EArr(np.where(np.asarray(_val(c)).astype(bool), np.asarray(_val(a)).astype(dt), np.asarray(_val(b)).astype(dt)))

    def minimum(self, a, b):
        dt = _target(a, b)
        if dt is None:                                   # jnp functions return (weak) float32 / int32 arrays
            dt = f32 if isinstance(a, float) or isinstance(b, float) else i32
        return EArr(np.minimum(np.asarray(_val(a)).astype(dt), np.asarray(_val(b)).astype(dt)))

    def maximum(self, a, b):
        dt = _target(a, b)
        if dt is None:
            dt = f32 if isinstance(a, float) or isinstance(b, float) else i32
        return EArr(np.maximum(np.asarray(_val(a)).astype(dt), np.asarray(_val(b)).astype(dt)))

    def clip(self, x, a_min=None, a_max=None):
        if a_min is not None:
            x = self.maximum(x, a_min)
        if a_max is not None:
            x = self.minimum(x, a_max)
        return x

    def _math(self, fn, x):
        with np.errstate(all="ignore"):
            return EArr(fn(np.asarray(_val(x)).astype(f32)).astype(f32))

    def abs(self, x): return abs(x) if _is_py(x) else EArr(np.abs(_arith(x.v)))
    def sqrt(self, x): return self._math(np.sqrt, x)
    def exp(self, x): return self._math(np.exp, x)
    def log(self, x): return self._math(np.log, x)
    def log1p(self, x): return self._math(np.log1p, x)
    def tanh(self, x): return self._math(np.tanh, x)
    def power(self, a, b): return _pow(a, b)
    def logical_and(self, a, b): return EArr(np.logical_and(_val(a), _val(b)))
    def logical_or(self, a, b): return EArr(np.logical_or(_val(a), _val(b)))
    def logical_not(self, a): return EArr(np.logical_not(_val(a)))

    def sum(self, x):
        v = x.v
        if v.dtype == np.dtype(f32):
            return EArr(_fsum(v))
        return EArr(i32(np.sum(v.astype(np.int64))))

    def mean(self, x):
        v = x.v
        s = _fsum(v) if v.dtype == np.dtype(f32) else f32(np.sum(v.astype(np.int64)))
        return EArr(f32(s / f32(v.shape[0])))

    def max(self, x): return EArr(np.max(x.v))
    def min(self, x): return EArr(np.min(x.v))
    def asarray(self, x, dtype=None): return EArr(x) if dtype is None else EArr(x).astype(dtype)
    array = asarray

    def nan_to_num(self, x, nan=0.0):
        v = np.asarray(_val(x)).astype(f32)
        return EArr(np.where(np.isnan(v), f32(nan), v))


jnp = _JNP()


class EKey:
    """A PRNG key or a column of keys (``(2,)`` or ``(n, 2)`` uint32) in the given stream layout."""

    def __init__(self, k, mode):
        self.k, self.mode = np.asarray(k, dtype=np.uint32), mode


class _Random:
    def split(self, key: EKey, num: int = 2):
        if key.k.ndim == 2:
            ch = jl.split_batched(key.k, num, key.mode)                  # [n, num, 2]
            return [EKey(ch[:, i], key.mode) for i in range(num)]
        ch = jl.split(key.k, num, key.mode)
        return [EKey(ch[i], key.mode) for i in range(num)]

    def _bits(self, key: EKey):
        if key.k.ndim == 2:
            return jl.random_bits_scalar_batched(key.k, key.mode)
        return jl.random_bits(key.k, (), key.mode)

    def uniform(self, key: EKey, shape=(), minval=0.0, maxval=1.0):
        return EArr(jl.bits_to_uniform(self._bits(key), minval, maxval))

    def normal(self, key: EKey, shape=()):
        return EArr(jl.bits_to_normal(self._bits(key)))


random = _Random()


class AgentTypeBase:
    """Stand-in for ``AgentType`` when user classes are built for this backend."""


def _unwrap_scalar(v):
    if isinstance(v, EArr):
        return v.v[()] if v.v.ndim == 0 else v.v
    return v


class wrap_agent_type:
    """Adapts a user ``AgentType`` (per-agent ``init_state`` / ``update`` bodies written against
    ``jnp`` / ``random``) to ``oracle.runtime``'s batch protocol: the body is evaluated once on whole
    columns with one key per agent -- what ``vmap`` does (``agent.py:125,173``)."""

    def __init__(self, user):
        self.user = user

    def init_batch(self, cfg, keys):
        from .runtime import unbatched
        out = self.user.init_state(cfg, EKey(keys, cfg.rng_mode))
        res = {}
        for k, v in out.items():
            if isinstance(v, EArr):
                res[k] = v.v if v.v.ndim >= 1 else unbatched(v.v[()])
            else:
                res[k] = v                      # Python scalar: broadcast by the runtime (vmap out_axes=0)
        return res

    def update_batch(self, s, model_state, cfg, keys):
        state = {k: EArr(v) for k, v in s.items()}
        env = {k: (EArr(v) if isinstance(v, np.generic) else v) for k, v in model_state["env"].items()}
        ms = {"time_step": model_state["time_step"], "env": env}
        out = self.user.update(state, ms, cfg, EKey(keys, cfg.rng_mode))
        n = keys.shape[0]
        res = {}
        for k, v in out.items():
            a = v.v if isinstance(v, EArr) else np.asarray(v)
            if a.ndim == 0:
                a = np.broadcast_to(a, (n,)).copy()
            res[k] = a.astype(s[k].dtype)
        return res


def wrap_model_fn(fn, mode, has_key=True):
    """update_state_fn / metrics_fn: env scalars that are NumPy scalars (results of earlier steps) enter
    as float32 / int32 values, Python scalars stay Python scalars -- exactly how the env's dtypes evolve
    in the reference, where nothing is jitted and every step is re-traced."""
    def wrapped(env, agent_states, params, key=None):
        e = {k: (EArr(v) if isinstance(v, np.generic) else v) for k, v in env.items()}
        a = {name: {f: EArr(col) for f, col in st.items()} for name, st in agent_states.items()}
        out = fn(e, a, params, EKey(key, mode)) if has_key else fn(e, a, params)
        return {k: _unwrap_scalar(v) for k, v in out.items()}
    return wrapped
