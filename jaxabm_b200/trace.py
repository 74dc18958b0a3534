"""Rule tracer: arbitrary user ``AgentType.update`` / ``update_state_fn`` / ``metrics_fn`` bodies ->
one fused sm_100a step kernel.

The reference evaluates user Python under ``jax.vmap`` (``jaxabm/agent.py:168-177``): JAX traces the
body into an expression graph and XLA compiles it.  This module does the same without JAX: the
body is run ONCE on symbolic values (:class:`Tr`) that record every operation with JAX's
x64-disabled dtype rules (weak Python scalars, ``bool < int32 < float32``), and the recorded graph is
emitted as CUDA C++ -- the per-agent update of every collection, the reductions that the model
functions ask for (``jnp.sum / mean / max / min`` over agent columns or expressions of them), and
the scalar env / metrics tail -- in the same single-launch shape as the hand-written
``step_kernel`` (``csrc/rules.cuh``).  ``nvcc`` compiles it for sm_100a into an in-tree shared
library that ``libjxb`` calls through two launcher entry points.

What can be traced: scalar (width-1) float32 / int32 / bool state fields; ``+ - * / ** //``-free
arithmetic, comparisons, ``& | ~``, ``where / minimum / maximum / clip / abs / sqrt / exp / log /
log1p / power / astype``; ``random.split / uniform / normal`` on the per-agent key and on the
update key; reads of ``model_state['env'][...]`` and ``model_state['time_step']``.  Python control
flow on traced values raises, exactly as it does under JAX.  Anything else raises
:class:`TraceError` -- there is still no CPU fallback.
"""
from __future__ import annotations

import math
import struct
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

F32, I32, BOOL, WF, WI = "f32", "i32", "bool", "wf64", "wi32"   # WF / WI: weak Python float / int scalars
_RANK = {BOOL: 0, WI: 1, I32: 1, WF: 2, F32: 2}


class TraceError(TypeError):
    """The rule uses something the tracer cannot turn into a kernel."""


_counter = [0]


def _next_id() -> int:
    _counter[0] += 1
    return _counter[0]


class Tr:
    """A traced value: an operation, its operands, a dtype and a scope ('a' per agent, 's' scalar)."""

    __array_priority__ = 1000
    __slots__ = ("op", "args", "dtype", "scope", "id", "attr")

    def __init__(self, op: str, args: tuple, dtype: str, scope: str, attr: Any = None):
        self.op, self.args, self.dtype, self.scope, self.attr = op, args, dtype, scope, attr
        self.id = _next_id()

    # -- Python protocol -------------------------------------------------------------------
    def __bool__(self):
        raise TraceError("the truth value of a traced array is not available at trace time (use jnp.where), "
                         "as under jax.vmap")

    def __float__(self):
        raise TraceError("float() of a traced value is not available at trace time")

    __int__ = __float__

    def __repr__(self):
        return f"Tr<{self.op}:{self.dtype}:{self.scope}#{self.id}>"

    @property
    def shape(self):
        return ()

    def astype(self, dt):
        return astype(self, dt)

    # arithmetic
    def __add__(self, o): return binary("add", self, o)
    def __radd__(self, o): return binary("add", o, self)
    def __sub__(self, o): return binary("sub", self, o)
    def __rsub__(self, o): return binary("sub", o, self)
    def __mul__(self, o): return binary("mul", self, o)
    def __rmul__(self, o): return binary("mul", o, self)
    def __truediv__(self, o): return binary("div", self, o)
    def __rtruediv__(self, o): return binary("div", o, self)
    def __pow__(self, o): return power(self, o)
    def __rpow__(self, o): return power(o, self)
    def __neg__(self): return unary("neg", self)
    def __pos__(self): return self
    def __abs__(self): return unary("abs", self)
    # comparisons
    def __lt__(self, o): return compare("lt", self, o)
    def __le__(self, o): return compare("le", self, o)
    def __gt__(self, o): return compare("gt", self, o)
    def __ge__(self, o): return compare("ge", self, o)
    def __eq__(self, o): return compare("eq", self, o)      # noqa: E711
    def __ne__(self, o): return compare("ne", self, o)
    __hash__ = object.__hash__
    # logic
    def __and__(self, o): return logical("and", self, o)
    def __rand__(self, o): return logical("and", o, self)
    def __or__(self, o): return logical("or", self, o)
    def __ror__(self, o): return logical("or", o, self)
    def __xor__(self, o): return logical("xor", self, o)
    def __invert__(self): return unary("not", self)


class TrKey:
    """A traced PRNG key: the per-agent key, the update key, or a child of ``split``."""
    __slots__ = ("kind", "parent", "index", "num", "id")

    def __init__(self, kind: str, parent: Optional["TrKey"] = None, index: int = 0, num: int = 0):
        self.kind, self.parent, self.index, self.num = kind, parent, index, num
        self.id = _next_id()

    @property
    def scope(self):
        k = self
        while k.parent is not None:
            k = k.parent
        return "a" if k.kind == "agent" else "s"


class TrVec:
    """A small fixed-length vector of traced scalars: a vector-valued state field (``f32[N, 2]`` positions), a
    non-scalar env entry, or ``jnp.array([...])`` of traced values.  Elementwise arithmetic with JAX's promotion
    rules per component; ``v[k]`` is component k.  ``n`` is set on the columns of a collection handed to the model
    functions (``len(positions)``)."""

    __array_priority__ = 1000

    def __init__(self, comps, n: Optional[int] = None):
        self.c = [(_const(x) if not isinstance(x, Tr) else x) for x in comps]
        self.n = n

    @property
    def shape(self):
        return (len(self.c),) if self.n is None else (self.n, len(self.c))

    @property
    def scope(self):
        return _scope(*self.c)

    def __len__(self):
        return len(self.c) if self.n is None else self.n

    def __iter__(self):
        if self.n is not None:
            raise TraceError("iterating over the agents of a traced column is not available at trace time")
        return iter(self.c)

    def __getitem__(self, k):
        if isinstance(k, tuple) and len(k) == 2 and k[0] in (slice(None), Ellipsis):
            k = k[1]
        if isinstance(k, (int, np.integer)):
            return self.c[int(k)]
        if isinstance(k, slice):
            return TrVec(self.c[k], self.n)
        raise TraceError("only v[k] / v[:, k] with a Python integer is traced on a vector value")

    def __array__(self, dtype=None, copy=None):
        if any(x.op != "const" for x in self.c):
            raise TraceError("a traced vector has no concrete value at trace time")
        dt = np.float32 if any(x.dtype in (F32, WF) for x in self.c) else (np.bool_ if all(x.dtype == BOOL for x in self.c) else np.int32)
        return np.array([x.attr for x in self.c], dtype=dtype or dt)

    def astype(self, dt):
        return TrVec([astype(x, dt) for x in self.c], self.n)

    def _zip(self, o, fn, rev=False):
        if isinstance(o, TrVec):
            if len(o.c) != len(self.c):
                raise TraceError(f"vector lengths differ: {len(self.c)} vs {len(o.c)}")
            os = o.c
        elif isinstance(o, (list, tuple, np.ndarray)) and np.ndim(o) == 1:
            if len(o) != len(self.c):
                raise TraceError(f"vector lengths differ: {len(self.c)} vs {len(o)}")
            os = [_const(x if not isinstance(o, np.ndarray) else o[i]) for i, x in enumerate(o)]
        else:
            os = [o] * len(self.c)
        n = self.n if self.n is not None else getattr(o, "n", None)
        return TrVec([fn(b, a) if rev else fn(a, b) for a, b in zip(self.c, os)], n)

    def __add__(self, o): return self._zip(o, lambda a, b: binary("add", a, b))
    def __radd__(self, o): return self._zip(o, lambda a, b: binary("add", a, b), True)
    def __sub__(self, o): return self._zip(o, lambda a, b: binary("sub", a, b))
    def __rsub__(self, o): return self._zip(o, lambda a, b: binary("sub", a, b), True)
    def __mul__(self, o): return self._zip(o, lambda a, b: binary("mul", a, b))
    def __rmul__(self, o): return self._zip(o, lambda a, b: binary("mul", a, b), True)
    def __truediv__(self, o): return self._zip(o, lambda a, b: binary("div", a, b))
    def __rtruediv__(self, o): return self._zip(o, lambda a, b: binary("div", a, b), True)
    def __pow__(self, o): return self._zip(o, lambda a, b: power(a, b))
    def __neg__(self): return TrVec([unary("neg", x) for x in self.c], self.n)
    def __abs__(self): return TrVec([unary("abs", x) for x in self.c], self.n)
    def __lt__(self, o): return self._zip(o, lambda a, b: compare("lt", a, b))
    def __le__(self, o): return self._zip(o, lambda a, b: compare("le", a, b))
    def __gt__(self, o): return self._zip(o, lambda a, b: compare("gt", a, b))
    def __ge__(self, o): return self._zip(o, lambda a, b: compare("ge", a, b))
    def __eq__(self, o): return self._zip(o, lambda a, b: compare("eq", a, b))      # noqa: E711
    def __ne__(self, o): return self._zip(o, lambda a, b: compare("ne", a, b))
    __hash__ = object.__hash__
    def __and__(self, o): return self._zip(o, lambda a, b: logical("and", a, b))
    def __or__(self, o): return self._zip(o, lambda a, b: logical("or", a, b))
    def __xor__(self, o): return self._zip(o, lambda a, b: logical("xor", a, b))
    def __invert__(self): return TrVec([unary("not", x) for x in self.c], self.n)

    def __bool__(self):
        raise TraceError("the truth value of a traced array is not available at trace time (use jnp.where)")


def make_vector(values) -> "TrVec":
    """``jnp.array([...])``: Python numbers become strongly typed float32 / int32 components (an all-int list is
    int32, any float makes it float32); traced components promote together the same way."""
    comps = [(_const(v) if not isinstance(v, Tr) else v) for v in values]
    kinds = [_RANK[c.dtype] for c in comps]
    top = max(kinds)
    target = {0: bool, 1: int, 2: float}[top]
    out = []
    for c in comps:
        if c.op == "const":
            out.append(Tr("const", (), {0: BOOL, 1: I32, 2: F32}[top], "s",
                          bool(c.attr) if top == 0 else (int(c.attr) if top == 1 else float(np.float32(c.attr)))))
        else:
            out.append(astype(c, target))
    return TrVec(out)


def _vmap(fn, *xs):
    """Apply a scalar tracer op elementwise when any operand is a vector."""
    vec = next((x for x in xs if isinstance(x, TrVec)), None)
    if vec is None:
        return fn(*xs)
    w = len(vec.c)
    cols = []
    for x in xs:
        if isinstance(x, TrVec):
            if len(x.c) != w:
                raise TraceError("vector lengths differ")
            cols.append(x.c)
        elif isinstance(x, (list, tuple, np.ndarray)) and np.ndim(x) == 1:
            cols.append([_const(v if not isinstance(x, np.ndarray) else x[i]) for i, v in enumerate(x)])
        else:
            cols.append([x] * w)
    n = next((x.n for x in xs if isinstance(x, TrVec) and x.n is not None), None)
    return TrVec([fn(*args) for args in zip(*cols)], n)


# ---------------------------------------------------------------------------------------------
# constructors
# ---------------------------------------------------------------------------------------------
def _const(v) -> Tr:
    if isinstance(v, Tr):
        return v
    if isinstance(v, (bool, np.bool_)):
        return Tr("const", (), BOOL, "s", bool(v))
    if isinstance(v, (int, np.integer)):
        if isinstance(v, np.integer):
            return Tr("const", (), I32, "s", int(v))
        return Tr("const", (), WI, "s", int(v))
    if isinstance(v, (float, np.floating)):
        if isinstance(v, np.float32):
            return Tr("const", (), F32, "s", float(v))
        return Tr("const", (), WF, "s", float(v))
    if isinstance(v, np.ndarray) and v.ndim == 0:
        return _const(v[()])
    raise TraceError(f"cannot trace a value of type {type(v).__name__} as a scalar")


def _scope(*xs: Tr) -> str:
    return "a" if any(x.scope == "a" for x in xs) else "s"


def _promote(a: Tr, b: Tr) -> str:
    """JAX's x64-disabled result dtype of a binary arithmetic op (weak scalars adopt the other side)."""
    da, db = a.dtype, b.dtype
    wa, wb = da in (WF, WI), db in (WF, WI)
    if wa and wb:
        return WF if WF in (da, db) else WI
    if wa or wb:
        weak, strong = (da, db) if wa else (db, da)
        if strong == F32:
            return F32
        if strong == I32:
            return F32 if weak == WF else I32
        return F32 if weak == WF else I32                      # bool (op) python scalar
    if F32 in (da, db):
        return F32
    if I32 in (da, db):
        return I32
    return BOOL


def _fold2(op: str, a: Tr, b: Tr):
    """Python-scalar (op) Python-scalar is evaluated by Python in double precision, as in the reference."""
    if a.op == "const" and b.op == "const" and a.dtype in (WF, WI) and b.dtype in (WF, WI):
        x, y = a.attr, b.attr
        if op == "div":
            return None if y == 0 else _const(x / y)
        if op in ("add", "sub", "mul"):
            return _const(x + y if op == "add" else (x - y if op == "sub" else x * y))
    return None


def _as_vec(x):
    """NumPy vectors (a non-scalar env entry such as 'bounds') take part in traced arithmetic as constant vectors."""
    if isinstance(x, np.ndarray) and x.ndim == 1:
        return TrVec([_const(v) for v in x])
    if isinstance(x, (list, tuple)) and x and all(isinstance(v, (int, float, np.number, Tr)) for v in x):
        return make_vector(x)
    return x


def binary(op: str, a, b) -> Tr:
    a, b = _as_vec(a), _as_vec(b)
    if isinstance(a, TrVec) or isinstance(b, TrVec):
        return _vmap(lambda x, y: binary(op, x, y), a, b)
    a, b = _const(a), _const(b)
    f = _fold2(op, a, b)
    if f is not None:
        return f
    dt = _promote(a, b)
    if op == "div":
        dt = WF if dt in (WI, WF) else F32                     # true division
    if dt == BOOL:
        dt = I32 if op in ("add", "sub", "mul") else dt
    return Tr(op, (a, b), dt, _scope(a, b))


def unary(op: str, a) -> Tr:
    if isinstance(a, TrVec):
        return TrVec([unary(op, x) for x in a.c], a.n)
    a = _const(a)
    if op == "not":
        if a.dtype != BOOL:
            raise TraceError("~ is only traced on boolean values")
        return Tr("not", (a,), BOOL, a.scope)
    if a.op == "const" and a.dtype in (WF, WI):
        return _const(-a.attr if op == "neg" else abs(a.attr))
    return Tr(op, (a,), I32 if a.dtype == BOOL else a.dtype, a.scope)


def compare(op: str, a, b) -> Tr:
    a, b = _as_vec(a), _as_vec(b)
    if isinstance(a, TrVec) or isinstance(b, TrVec):
        return _vmap(lambda x, y: compare(op, x, y), a, b)
    a, b = _const(a), _const(b)
    return Tr(op, (a, b), BOOL, _scope(a, b), _promote(a, b))


def logical(op: str, a, b) -> Tr:
    if isinstance(a, TrVec) or isinstance(b, TrVec):
        return _vmap(lambda x, y: logical(op, x, y), a, b)
    a, b = _const(a), _const(b)
    if a.dtype != BOOL or b.dtype != BOOL:
        raise TraceError("& | ^ are only traced on boolean values")
    return Tr(op, (a, b), BOOL, _scope(a, b))


def power(a, b) -> Tr:
    if isinstance(a, TrVec) or isinstance(b, TrVec):
        return _vmap(power, a, b)
    a, b = _const(a), _const(b)
    if b.op == "const" and b.dtype == WI and 0 <= b.attr <= 64:
        # lax.integer_pow: square-and-multiply (x**4 = (x*x)*(x*x), x**3 = (x*x)*x ... the order decides the
        # float32 rounding); x**0 is the constant 1 of the operand's dtype, also for inf / nan inputs
        n = int(b.attr)
        if n == 0:
            dt = a.dtype if a.dtype in (F32, WF, I32, WI) else I32
            return Tr("const", (), dt, "s", 1.0 if dt in (F32, WF) else 1)
        acc, base = None, a
        while n > 0:
            if n & 1:
                acc = base if acc is None else binary("mul", acc, base)
            n >>= 1
            if n > 0:
                base = binary("mul", base, base)
        return acc
    if a.op == "const" and b.op == "const" and a.dtype in (WF, WI) and b.dtype in (WF, WI):
        return _const(float(a.attr) ** float(b.attr))
    dt = _promote(a, b)
    return Tr("pow", (a, b), WF if dt in (WF, WI) else F32, _scope(a, b))


def _math(op: str, a) -> Tr:
    if isinstance(a, TrVec):
        return TrVec([_math(op, x) for x in a.c], a.n)
    a = _const(a)
    return Tr(op, (a,), F32, a.scope)


def _array_dtype(dt: str) -> str:
    """jnp functions return arrays: a Python-scalar result becomes (weak) float32 / int32, it does not
    stay a double-precision Python number."""
    return F32 if dt == WF else (I32 if dt == WI else dt)


def where(c, a, b) -> Tr:
    a, b = _as_vec(a), _as_vec(b)
    if isinstance(c, TrVec) or isinstance(a, TrVec) or isinstance(b, TrVec):
        return _vmap(where, c, a, b)
    c, a, b = _const(c), _const(a), _const(b)
    if c.dtype != BOOL:
        c = compare("ne", c, 0)
    return Tr("where", (c, a, b), _array_dtype(_promote(a, b)), _scope(c, a, b))


def minimum(a, b) -> Tr:
    a, b = _as_vec(a), _as_vec(b)
    if isinstance(a, TrVec) or isinstance(b, TrVec):
        return _vmap(minimum, a, b)
    a, b = _const(a), _const(b)
    return Tr("min", (a, b), _array_dtype(_promote(a, b)), _scope(a, b))


def maximum(a, b) -> Tr:
    a, b = _as_vec(a), _as_vec(b)
    if isinstance(a, TrVec) or isinstance(b, TrVec):
        return _vmap(maximum, a, b)
    a, b = _const(a), _const(b)
    return Tr("max", (a, b), _array_dtype(_promote(a, b)), _scope(a, b))


def clip(x, lo, hi) -> Tr:
    return minimum(maximum(x, lo), hi)


def astype(a, dt) -> Tr:
    if isinstance(a, TrVec):
        return a.astype(dt)
    a = _const(a)
    name = getattr(dt, "__name__", str(dt))
    if dt in (float, np.float32, np.float64) or name in ("float", "float32", "float64"):
        target = F32
    elif dt in (int, np.int32, np.int64) or name in ("int", "int32", "int64"):
        target = I32
    elif dt in (bool, np.bool_) or name in ("bool", "bool_"):
        target = BOOL
    else:
        raise TraceError(f"astype({dt!r}) is not traced")
    if a.dtype == target:
        return a
    return Tr("cast", (a,), target, a.scope)


class Column:
    """``agent_states[name][field]`` inside update_state_fn / metrics_fn: the POST-update column of a
    collection, usable inside ``jnp.sum / mean / max / min`` (alone or inside an expression)."""

    def __init__(self, type_index: int, field_index: int, dtype: str, n: int):
        self.tr = Tr("field", (), dtype, "a", (type_index, field_index))
        self.n = n


class TrCol(Tr):
    """A scalar state column of a collection as the model functions see it (``agent_states[name][field]``):
    a per-agent value that also knows the population size (``len(column)``)."""
    __slots__ = ("n",)

    def __len__(self):
        return self.n


def _fold(kind: str, comps):
    """sum / max / min / mean over the components of a vector, left to right."""
    acc = comps[0]
    for c in comps[1:]:
        acc = binary("add", acc, c) if kind in ("sum", "mean") else (maximum(acc, c) if kind == "max" else minimum(acc, c))
    if kind == "mean":
        acc = binary("div", astype(acc, float), float(len(comps)))
    return acc


def _reduce(kind: str, x, axis=None):
    if isinstance(x, Column):
        x = x.tr
    x = _as_vec(x)
    if isinstance(x, TrVec):
        is_matrix = x.n is not None                      # (N, w): the columns of a collection
        if axis is not None and axis < 0:
            axis += 2 if is_matrix else 1
        if not is_matrix:
            if axis not in (None, 0):
                raise TraceError(f"axis {axis} is out of range for a vector")
            return _fold(kind, x.c)
        if axis == 1:
            return _fold(kind, x.c)                       # per agent, over the components
        if axis == 0:
            return TrVec([_reduce(kind, c) for c in x.c])  # per component, over the agents
        if kind == "mean":
            return binary("div", _reduce("sum", _fold("sum", x.c)), float(x.n * len(x.c)))
        return _reduce(kind, _fold(kind, x.c))
    if axis not in (None, 0):
        raise TraceError(f"axis {axis} is out of range for a column")
    x = _const(x)
    if x.scope != "a":
        return x
    types = _types_of(x)
    if len(types) != 1:
        raise TraceError("a reduction must range over exactly one agent collection")
    dt = x.dtype
    if kind == "mean":
        dt = F32
    elif dt == BOOL:
        dt = I32 if kind == "sum" else BOOL
    return Tr("reduce", (x,), dt, "s", (kind, types.pop()))


def _types_of(x: Tr, seen=None) -> set:
    out = set()
    stack = [x]
    visited = set()
    while stack:
        v = stack.pop()
        if v.id in visited:
            continue
        visited.add(v.id)
        if v.op in ("field", "field_new"):
            out.add(v.attr[0])
        stack.extend(a for a in v.args if isinstance(a, Tr))
    return out


# ---------------------------------------------------------------------------------------------
# the jnp-like namespace handed to user code (jaxabm_b200.numpy)
# ---------------------------------------------------------------------------------------------
class _Namespace:
    float32, int32, bool_ = np.float32, np.int32, np.bool_
    pi, e, inf, nan = math.pi, math.e, math.inf, math.nan

    @staticmethod
    def _col(x):
        return x.tr if isinstance(x, Column) else x

    def where(self, c, a, b): return where(self._col(c), self._col(a), self._col(b))
    def minimum(self, a, b): return minimum(self._col(a), self._col(b))
    def maximum(self, a, b): return maximum(self._col(a), self._col(b))
    def clip(self, x, a_min=None, a_max=None):
        x = self._col(x)
        if a_min is not None:
            x = maximum(x, a_min)
        if a_max is not None:
            x = minimum(x, a_max)
        return x
    def abs(self, x): return unary("abs", self._col(x))
    absolute = abs
    def sqrt(self, x): return _math("sqrt", self._col(x))
    def exp(self, x): return _math("exp", self._col(x))
    def log(self, x): return _math("log", self._col(x))
    def log1p(self, x): return _math("log1p", self._col(x))
    def tanh(self, x): return _math("tanh", self._col(x))
    def power(self, a, b): return power(self._col(a), self._col(b))
    def logical_and(self, a, b): return logical("and", self._col(a), self._col(b))
    def logical_or(self, a, b): return logical("or", self._col(a), self._col(b))
    def logical_not(self, a): return unary("not", self._col(a))
    def sum(self, x, axis=None): return _reduce("sum", x, axis)
    def mean(self, x, axis=None): return _reduce("mean", x, axis)
    def max(self, x, axis=None): return _reduce("max", x, axis)
    def min(self, x, axis=None): return _reduce("min", x, axis)
    def asarray(self, x, dtype=None):
        x = self._col(x)
        if isinstance(x, TrVec):
            return x.astype(dtype) if dtype is not None else x
        if isinstance(x, (list, tuple)) or (isinstance(x, np.ndarray) and x.ndim == 1):
            v = make_vector(list(x))
            return v.astype(dtype) if dtype is not None else v
        return astype(_const(x), dtype) if dtype is not None else _const(x)
    array = asarray
    def square(self, x): return binary("mul", self._col(x), self._col(x))
    def nan_to_num(self, x, nan=0.0):
        x = _const(self._col(x))
        return where(compare("ne", x, x), nan, x)


numpy = _Namespace()


# ---------------------------------------------------------------------------------------------
# traced jax.random
# ---------------------------------------------------------------------------------------------
def key_split(key: TrKey, num: int = 2) -> List[TrKey]:
    return [TrKey("split", key, i, int(num)) for i in range(int(num))]


def key_uniform(key: TrKey, shape=(), minval=0.0, maxval=1.0) -> Tr:
    if shape not in ((), None):
        raise TraceError("only scalar random draws are traced")
    if isinstance(minval, Tr) or isinstance(maxval, Tr):
        raise TraceError("uniform(minval, maxval) bounds must be Python numbers")
    return Tr("uniform", (), F32, key.scope, (key, float(minval), float(maxval)))


def key_normal(key: TrKey, shape=()) -> Tr:
    if shape not in ((), None):
        raise TraceError("only scalar random draws are traced")
    return Tr("normal", (), F32, key.scope, (key,))


# ---------------------------------------------------------------------------------------------
# code generation
# ---------------------------------------------------------------------------------------------
def _f32_lit(v: float) -> str:
    if v != v:
        return "__int_as_float(0x7fc00000)"
    if v in (math.inf, -math.inf):
        return ("" if v > 0 else "-") + "__int_as_float(0x7f800000)"
    f = struct.unpack("f", struct.pack("f", v))[0]
    return f"{f:.9g}f" if ("." in f"{f:.9g}" or "e" in f"{f:.9g}" or "n" in f"{f:.9g}") else f"{f:.9g}.0f"


def _f64_lit(v: float) -> str:
    if v != v:
        return "(0.0/0.0)"
    if v in (math.inf, -math.inf):
        return ("" if v > 0 else "-") + "(1.0/0.0)"
    s = repr(float(v))
    return s if ("." in s or "e" in s) else s + ".0"


_CT = {F32: "float", I32: "int", BOOL: "bool", WF: "double", WI: "int"}


_SEQ = [0]


class Emitter:
    """Emits the statements of an expression DAG once each (common sub-expressions are shared)."""

    def __init__(self, leaf, mode_key: str, pool: Optional[List[float]] = None):
        self.pool = pool                   # float constants go to a table (values) instead of into the code
        self.seq = _SEQ                    # statement names are numbered in emission order, not by node id, so
        self.lines: List[str] = []         # that two traces of the same code give byte-identical sources
        self.names: Dict[int, str] = {}
        self.key_names: Dict[int, str] = {}
        self.leaf = leaf                   # callable(Tr) -> C expression for 'field' / 'env' / 'time' / 'reduce' leaves
        self.mode_key = mode_key           # C expression of the root key for this scope

    def cast(self, x: Tr, to: str) -> str:
        s = self.ref(x)
        if x.dtype == to or (x.dtype == WI and to == I32) or (x.dtype == I32 and to == WI):
            return s
        return f"(({_CT[to]}){s})"

    def key(self, k: TrKey) -> str:
        if k.id in self.key_names:
            return self.key_names[k.id]
        if k.parent is None:
            name = self.mode_key
        else:
            self.seq[0] += 1
            name = f"k{self.seq[0]}"
            self.lines.append(f"const Key {name} = split_child<MODE>({self.key(k.parent)}, {k.index}ull, {k.num}ull);")
        self.key_names[k.id] = name
        return name

    def ref(self, x: Tr) -> str:
        if x.id in self.names:
            return self.names[x.id]
        expr = self.expr(x)
        if x.op == "const":
            self.names[x.id] = expr
            return expr
        self.seq[0] += 1
        name = f"v{self.seq[0]}"
        self.lines.append(f"const {_CT[x.dtype]} {name} = {expr};")
        self.names[x.id] = name
        return name

    def expr(self, x: Tr) -> str:
        op, dt = x.op, x.dtype
        if op == "const":
            if dt == BOOL:
                return "true" if x.attr else "false"
            if dt in (I32, WI):
                return f"{int(x.attr)}"
            if self.pool is not None:
                self.pool.append(float(x.attr))
                return f"(float)cst[{len(self.pool) - 1}]" if dt == F32 else f"cst[{len(self.pool) - 1}]"
            return _f32_lit(x.attr) if dt == F32 else _f64_lit(x.attr)
        if op in ("field", "field_new", "env", "time", "reduce"):
            return self.leaf(x, self)
        if op in ("add", "sub", "mul", "div"):
            a, b = self.cast(x.args[0], dt), self.cast(x.args[1], dt)
            return f"{a} {'+-*/'['add sub mul div'.split().index(op)]} {b}"
        if op == "neg":
            return f"-{self.cast(x.args[0], dt)}"
        if op == "abs":
            a = self.cast(x.args[0], dt)
            return f"fabsf({a})" if dt == F32 else (f"fabs({a})" if dt == WF else f"abs({a})")
        if op in ("lt", "le", "gt", "ge", "eq", "ne"):
            ct = x.attr if x.attr != BOOL else I32
            a, b = self.cast(x.args[0], ct), self.cast(x.args[1], ct)
            return f"{a} {dict(lt='<', le='<=', gt='>', ge='>=', eq='==', ne='!=')[op]} {b}"
        if op in ("and", "or", "xor"):
            return f"{self.ref(x.args[0])} {dict(**{'and': '&&', 'or': '||', 'xor': '!='})[op]} {self.ref(x.args[1])}"
        if op == "not":
            return f"!{self.ref(x.args[0])}"
        if op == "where":
            return f"{self.ref(x.args[0])} ? {self.cast(x.args[1], dt)} : {self.cast(x.args[2], dt)}"
        if op in ("min", "max"):
            a, b = self.cast(x.args[0], dt), self.cast(x.args[1], dt)
            if dt == F32:
                return f"{'jmin' if op == 'min' else 'jmax'}({a}, {b})"
            if dt == WF:
                return f"(({a} != {a} || {b} != {b}) ? (0.0/0.0) : f{op}({a}, {b}))"
            return f"{op}({a}, {b})"
        if op == "pow":
            return f"powf({self.cast(x.args[0], F32)}, {self.cast(x.args[1], F32)})" if dt == F32 else \
                f"pow({self.cast(x.args[0], WF)}, {self.cast(x.args[1], WF)})"
        if op in ("sqrt", "exp", "log", "log1p", "tanh"):
            return f"{op}f({self.cast(x.args[0], F32)})"
        if op == "cast":
            a = x.args[0]
            if dt == BOOL:
                return f"{self.ref(a)} != 0"
            return f"({_CT[dt]}){self.ref(a)}"
        if op == "uniform":
            k, lo, hi = x.attr
            return f"bits_to_uniform(bits_scalar<MODE>({self.key(k)}), {_f32_lit(lo)}, {_f32_lit(hi)})"
        if op == "normal":
            return f"normal_scalar<MODE>({self.key(x.attr[0])})"
        raise TraceError(f"no code generation for traced op {op!r}")


def used_leaves(roots: Sequence[Tr], op: str) -> List[Tr]:
    out, seen, stack = {}, set(), list(roots)
    while stack:
        v = stack.pop()
        if v.id in seen:
            continue
        seen.add(v.id)
        if v.op == op:
            out[v.id] = v
        stack.extend(a for a in v.args if isinstance(a, Tr))
    return list(out.values())


# ---------------------------------------------------------------------------------------------
# tracing a whole model and emitting its step kernel
# ---------------------------------------------------------------------------------------------
_DT_CODE = {F32: 0, I32: 1, BOOL: 2, WF: 3, WI: 1}
_STORE = {F32: F32, WF: F32, I32: I32, WI: I32, BOOL: BOOL}


def _env_dtype_of(v) -> Optional[str]:
    if isinstance(v, Tr):
        return v.dtype
    if isinstance(v, (bool, np.bool_)):
        return BOOL
    if isinstance(v, (int, np.integer)):
        return WI if isinstance(v, int) else I32
    if isinstance(v, np.float32):
        return F32
    if isinstance(v, (float, np.floating)):
        return WF
    if isinstance(v, np.ndarray) and v.ndim == 0:
        return _env_dtype_of(v[()])
    return None


class TracedModel:
    """Everything the code generator needs, for one env dtype signature (variant)."""

    def __init__(self):
        self.types: List[dict] = []        # {name, n, fields [(name, dtype)], init {field: Tr}, update {field: Tr}}
        self.env_names: List[str] = []
        self.env_dtypes: List[str] = []
        self.env_out: Dict[str, Tr] = {}
        self.metrics: List[Tuple[str, Tr]] = []
        self.has_env_fn = False


def trace_model(collections, env_state: dict, params: dict, update_state_fn, metrics_fn, config,
                env_dtypes: Optional[Dict[str, str]] = None, slot_order: Optional[List[str]] = None) -> TracedModel:
    """Run every user function once on symbolic values.  ``env_dtypes``: dtype of every scalar env
    entry at the START of the traced step (None: the Python types of ``env_state``); entries named
    there but absent from ``env_state`` are the ones an earlier step's update_state_fn added."""
    tm = TracedModel()
    scalars = {k: v for k, v in env_state.items() if _env_dtype_of(v) is not None}
    dts = {k: _env_dtype_of(v) for k, v in scalars.items()}
    visible = list(scalars)
    if env_dtypes:
        for k, v in env_dtypes.items():
            if k not in dts:
                visible.append(k)
            dts[k] = v
    tm.env_names = list(slot_order) if slot_order else list(visible)

    def env_view():
        d = dict(env_state)
        for name in visible:
            d[name] = Tr("env", (), dts[name], "s", tm.env_names.index(name))
        return d

    for ti, (cname, coll) in enumerate(collections.items()):
        at = coll.agent_type
        init = at.init_state(config, TrKey("agent"))
        if not isinstance(init, dict) or not init:
            raise TraceError(f"{type(at).__name__}.init_state must return a non-empty dict")
        # fields: (name, dtype, width); init / update hold one traced scalar per component
        fields, init_tr = [], {}
        for fname, v in init.items():
            v = _as_vec(v)
            comps = list(v.c) if isinstance(v, TrVec) else [(_const(v) if not isinstance(v, Tr) else v)]
            if not 1 <= len(comps) <= 4:
                raise TraceError(f"state field {fname!r}: vector fields of 1..4 components are traced, got {len(comps)}")
            dt = _STORE[max((c.dtype for c in comps), key=lambda d: _RANK[d])]
            fields.append((fname, dt, len(comps)))
            init_tr[fname] = [c if _STORE[c.dtype] == dt else astype(c, {F32: float, I32: int, BOOL: bool}[dt]) for c in comps]

        def column(op, fi, dt, w, n=None, ti=ti):
            if w == 1:
                if n is None:
                    return Tr(op, (), dt, "a", (ti, fi, 0))
                col = TrCol(op, (), dt, "a", (ti, fi, 0))
                col.n = n
                return col
            return TrVec([Tr(op, (), dt, "a", (ti, fi, c)) for c in range(w)], n)
        state = {fname: column("field", fi, dt, w) for fi, (fname, dt, w) in enumerate(fields)}
        model_state = {"time_step": Tr("time", (), WI, "s"), "env": env_view()}
        out = at.update(dict(state), model_state, config, TrKey("agent"))
        if not isinstance(out, dict) or set(out) != set(state):
            raise TraceError(f"{type(at).__name__}.update must return the same state keys as init_state "
                             f"({sorted(state)}), got {sorted(out) if isinstance(out, dict) else type(out).__name__}")
        upd = {}
        for fname, dt, w in fields:
            v = _as_vec(out[fname])
            comps = list(v.c) if isinstance(v, TrVec) else [(_const(v) if not isinstance(v, Tr) else v)]
            if len(comps) != w:
                raise TraceError(f"update returns {len(comps)} components for state field {fname!r} of width {w}")
            upd[fname] = [c if _STORE[c.dtype] == dt else astype(c, {F32: float, I32: int, BOOL: bool}[dt])   # the column keeps its dtype
                          for c in comps]
        tm.types.append({"name": cname, "n": coll.num_agents, "fields": fields, "init": init_tr, "update": upd,
                         "state": state, "column": column})
    # agent_states[name][field]: the POST-update column of a collection (model.py:182-200 hands the new
    # states to update_state_fn / metrics_fn); usable inside jnp.sum / mean / max / min
    agent_states = {t["name"]: {fname: t["column"]("field_new", fi, dt, w, t["n"]) for fi, (fname, dt, w) in enumerate(t["fields"])}
                    for t in tm.types}
    env_after = env_view()
    if update_state_fn is not None:
        tm.has_env_fn = True
        new_env = update_state_fn(dict(env_after), agent_states, params, TrKey("update"))
        if not isinstance(new_env, dict):
            raise TraceError("update_state_fn must return the env dict")
        for k, v in new_env.items():
            old = env_after.get(k)
            if v is old:
                continue
            if _env_dtype_of(v) is None:
                if isinstance(old, Tr):
                    raise TraceError(f"env entry {k!r} becomes non-scalar")
                continue
            if k not in tm.env_names:
                tm.env_names.append(k)
                dts[k] = _env_dtype_of(v)
            tm.env_out[k] = _const(v) if not isinstance(v, Tr) else v
        env_after = dict(new_env)
        for name in visible:                                   # entries the fn dropped keep their leaf
            env_after.setdefault(name, Tr("env", (), dts[name], "s", tm.env_names.index(name)))
    if metrics_fn is not None:
        met = metrics_fn(env_after, agent_states, params)
        if not isinstance(met, dict):
            raise TraceError("metrics_fn must return a dict")
        for k, v in met.items():
            if _env_dtype_of(v) is None:
                raise TraceError(f"metric {k!r} is not a scalar")
            tm.metrics.append((k, _const(v) if not isinstance(v, Tr) else v))
    tm.env_dtypes = [dts.get(n, WF) for n in tm.env_names]
    return tm


def _acc_native(dt: str) -> str:
    return {F32: "float", WF: "float", I32: "int", WI: "int", BOOL: "int"}[dt]


def generate_source(variants: List[TracedModel]) -> Tuple[str, dict]:
    """CUDA C++ of the step kernel (one template instance per env-dtype variant) + the init kernel."""
    tm0 = variants[0]
    n_types = len(tm0.types)
    out: List[str] = []
    w = out.append
    _SEQ[0] = 0
    pool: List[float] = []               # float constants, in emission order (position-based: the SOURCE does
                                         # not depend on their values, so a parameter sweep compiles once)
    w('// generated by jaxabm_b200/trace.py -- do not edit\n#include <string.h>\n#include "common.cuh"\n#include "economy.cuh"\nusing namespace jxb;\n')
    # reductions are numbered per variant (different dtype signatures may trace different graphs)
    meta = {"n_acc": 0, "n_variants": len(variants)}
    for var, tm in enumerate(variants):
        roots = list(tm.env_out.values()) + [v for _, v in tm.metrics]
        reds = used_leaves(roots, "reduce")
        meta["n_acc"] = max(meta["n_acc"], len(reds))
        red_slot = {r.id: i for i, r in enumerate(reds)}
        nacc = max(len(reds), 1)
        # ---- per-type agent code --------------------------------------------------------------
        # jxc_one_*: the traced update of ONE agent (inputs by value, new columns and accumulators by
        # reference); jxc_agents_*: the streaming driver -- four agents per thread per iteration with
        # one 16-byte (4-byte for bool) load / store per column, like the hand-written kernels.
        for ti, t in enumerate(tm.types):
            my_reds = [r for r in reds if r.attr[1] == ti]
            fields = t["fields"]                      # (name, dtype, width)

            def leaf(x, em, ti=ti):
                if x.op == "field":
                    tj, fj, cj = x.attr
                    if tj != ti:
                        raise TraceError("reading another collection's state inside update is not traced")
                    return f"f{fj}_{cj}"
                if x.op == "env":
                    return _env_load(x, "env")
                if x.op == "time":
                    return "(int)time_step"
                raise TraceError("reductions cannot be used inside a per-agent update")
            em = Emitter(leaf, "ak", pool)
            upd_roots = [v for fname, _, _ in fields for v in t["update"][fname]]
            used = {x.attr[1] for x in used_leaves(upd_roots + [r.args[0] for r in my_reds], "field") if x.attr[0] == ti}
            new_names = {}                            # (field, component) -> expression of the new value, None = unchanged
            for fi, (fname, dt, wd) in enumerate(fields):
                for c in range(wd):
                    v = t["update"][fname][c]
                    new_names[(fi, c)] = em.cast(v, dt) if v.op != "field" or tuple(v.attr) != (ti, fi, c) else None

            def leaf_new(x, em2, ti=ti, new_names=new_names):          # reductions read the NEW columns
                if x.op == "field_new":
                    tj, fj, cj = x.attr
                    return new_names[(fj, cj)] if new_names[(fj, cj)] is not None else f"f{fj}_{cj}"
                if x.op == "env":
                    return _env_load(x, "env")
                if x.op == "time":
                    return "(int)time_step"
                raise TraceError("unsupported leaf inside a reduction")
            em2 = Emitter(leaf_new, "ak", pool)
            em2.names, em2.lines, em2.key_names = em.names, em.lines, em.key_names     # share the statement list
            red_exprs = {}
            for r in my_reds:
                red_exprs[r.id] = em2.cast(r.args[0], {"float": F32, "int": I32}[_acc_native(r.args[0].dtype)])
                for x in used_leaves([r.args[0]], "field_new"):
                    if new_names[(x.attr[1], x.attr[2])] is None:
                        used.add(x.attr[1])
            stored = [fi for fi, (_, _, wd) in enumerate(fields) if any(new_names[(fi, c)] is not None for c in range(wd))]
            for fi in stored:                         # an unchanged component of a rewritten vector field is copied through
                if any(new_names[(fi, c)] is None for c in range(fields[fi][2])):
                    used.add(fi)
            loaded = [fi for fi in range(len(fields)) if fi in used]
            needs_key = any("ak" in ln for ln in em.lines)
            sig_in = "".join(f", const {_CT[fields[fi][1]]} f{fi}_{c}" for fi in loaded for c in range(fields[fi][2]))
            sig_out = "".join(f", {_CT[fields[fi][1]]}& n{fi}_{c}" for fi in stored for c in range(fields[fi][2]))
            sig_acc = "".join(f", {_acc_native(r.args[0].dtype)}& a{red_slot[r.id]}" for r in my_reds)
            w(f"template <int MODE> __device__ __forceinline__ void jxc_one_v{var}_t{ti}(const TypeDev& t, const double* env, "
              f"const double* __restrict__ cst, long long time_step, Key ck, long long i{sig_in}{sig_out}{sig_acc}) {{")
            if needs_key:
                w("  const Key ak = split_child<MODE>(ck, (unsigned long long)(t.goff + i), (unsigned long long)t.gn);")
            for ln in em.lines:
                w("  " + ln)
            for fi in stored:
                for c in range(fields[fi][2]):
                    w(f"  n{fi}_{c} = {new_names[(fi, c)] if new_names[(fi, c)] is not None else f'f{fi}_{c}'};")
            for r in my_reds:
                sl, kind, nat = red_slot[r.id], r.attr[0], _acc_native(r.args[0].dtype)
                e = red_exprs[r.id]
                if kind in ("sum", "mean"):
                    w(f"  a{sl} += {e};")
                elif nat == "float":
                    w(f"  a{sl} = {'jmax' if kind == 'max' else 'jmin'}(a{sl}, {e});")
                else:
                    w(f"  a{sl} = {kind}(a{sl}, {e});")
            w("}\n")
            # ---- driver ------------------------------------------------------------------------
            # four agents per thread per iteration; a column of width w is w 16-byte (4-byte for bool) vectors per
            # group: component c of agent j of the group is flat element j*w + c
            VT = {F32: ("float4", "float"), I32: ("int4", "int"), BOOL: ("uchar4", "unsigned char")}

            def elem(prefix, fi, j, c):
                flat = j * fields[fi][2] + c
                return f"{prefix}{fi}_{flat // 4}.{'xyzw'[flat % 4]}"
            w(f"template <int MODE> __device__ __forceinline__ void jxc_agents_v{var}_t{ti}(const TypeDev& t, const double* env, "
              f"const double* __restrict__ cst, long long time_step, Key ck, int lb, double* accd) {{")
            for r in my_reds:
                kind, nat = r.attr[0], _acc_native(r.args[0].dtype)
                init = "0" if kind in ("sum", "mean") else (
                    ("-__int_as_float(0x7f800000)" if kind == "max" else "__int_as_float(0x7f800000)") if nat == "float"
                    else ("-2147483647-1" if kind == "max" else "2147483647"))
                w(f"  {nat} a{red_slot[r.id]} = {init};")
            acc_args = "".join(f", a{red_slot[r.id]}" for r in my_reds)
            w("  const long long ngroups = t.n / 4, stride = (long long)t.block_count * blockDim.x;")
            w("  for (long long g = (long long)lb * blockDim.x + threadIdx.x; g < ngroups; g += stride) {")
            for fi in loaded:
                vt, _ = VT[fields[fi][1]]
                for v in range(fields[fi][2]):
                    w(f"    const {vt} F{fi}_{v} = *(({vt}*)t.f[{fi}] + g * {fields[fi][2]} + {v});")
            for fi in stored:
                vt, _ = VT[fields[fi][1]]
                for v in range(fields[fi][2]):
                    w(f"    {vt} N{fi}_{v};")
            for j in range(4):
                ins = "".join(f", {elem('F', fi, j, c)}{' != 0' if fields[fi][1] == BOOL else ''}"
                              for fi in loaded for c in range(fields[fi][2]))
                for fi in stored:
                    for c in range(fields[fi][2]):
                        w(f"    {_CT[fields[fi][1]]} n{fi}_{c}_{j};")
                outs = "".join(f", n{fi}_{c}_{j}" for fi in stored for c in range(fields[fi][2]))
                w(f"    jxc_one_v{var}_t{ti}<MODE>(t, env, cst, time_step, ck, 4 * g + {j}{ins}{outs}{acc_args});")
                for fi in stored:
                    for c in range(fields[fi][2]):
                        w(f"    {elem('N', fi, j, c)} = n{fi}_{c}_{j}{' ? 1 : 0' if fields[fi][1] == BOOL else ''};")
            for fi in stored:
                vt, _ = VT[fields[fi][1]]
                for v in range(fields[fi][2]):
                    w(f"    *(({vt}*)t.f[{fi}] + g * {fields[fi][2]} + {v}) = N{fi}_{v};")
            w("  }")
            w("  for (long long i = ngroups * 4 + threadIdx.x; lb == 0 && i < t.n; i += blockDim.x) {   // tail agents")
            for fi in loaded:
                _, st = VT[fields[fi][1]]
                for c in range(fields[fi][2]):
                    w(f"    const {_CT[fields[fi][1]]} f{fi}_{c} = (({st}*)t.f[{fi}])[i * {fields[fi][2]} + {c}]{' != 0' if fields[fi][1] == BOOL else ''};")
            for fi in stored:
                for c in range(fields[fi][2]):
                    w(f"    {_CT[fields[fi][1]]} n{fi}_{c};")
            ins = "".join(f", f{fi}_{c}" for fi in loaded for c in range(fields[fi][2]))
            outs = "".join(f", n{fi}_{c}" for fi in stored for c in range(fields[fi][2]))
            w(f"    jxc_one_v{var}_t{ti}<MODE>(t, env, cst, time_step, ck, i{ins}{outs}{acc_args});")
            for fi in stored:
                _, st = VT[fields[fi][1]]
                for c in range(fields[fi][2]):
                    w(f"    (({st}*)t.f[{fi}])[i * {fields[fi][2]} + {c}] = n{fi}_{c}{' ? 1 : 0' if fields[fi][1] == BOOL else ''};")
            w("  }")
            for r in my_reds:
                sl, kind, nat = red_slot[r.id], r.attr[0], _acc_native(r.args[0].dtype)
                if kind in ("sum", "mean"):
                    w(f"  accd[{sl}] = (double)warp_sum(a{sl});")
                elif nat == "float":
                    w(f"  accd[{sl}] = (double)warp_{kind}(a{sl});" if kind == "max" else f"  accd[{sl}] = -(double)warp_max(-a{sl});")
                else:
                    w(f"  {{ int v = a{sl}; for (int o = 16; o > 0; o >>= 1) v = {kind}(v, __shfl_xor_sync(0xffffffffu, v, o)); accd[{sl}] = (double)v; }}")
            w("}\n")
        # ---- tail -----------------------------------------------------------------------------
        w(f"template <int MODE> __device__ inline void jxc_tail_v{var}(const ModelDev& md, const double* tot, Key uk, double* m) {{")
        w("  double* env = md.env;\n  const double* __restrict__ cst = md.consts;\n  const long long time_step = md.ctrl->time_step;")

        def leaf_tail(x, em, tm=tm, red_slot=red_slot):
            if x.op == "env":
                return _env_load(x, "env")
            if x.op == "time":
                return "(int)time_step"
            if x.op == "reduce":
                kind, ti = x.attr
                s = red_slot[x.id]
                nat = _acc_native(x.args[0].dtype)
                if kind == "mean":
                    return f"((float)tot[{s}] / (float)md.t[{ti}].gn)"
                if nat == "float":
                    return f"(float)tot[{s}]"
                return f"(int)tot[{s}]" if x.dtype != BOOL else f"(tot[{s}] != 0.0)"
            raise TraceError("agent columns can only be used inside jnp.sum / mean / max / min in model functions")
        em = Emitter(leaf_tail, "uk", pool)
        env_exprs = {k: em.ref(v) for k, v in tm.env_out.items()}
        met_exprs = [(k, em.ref(v)) for k, v in tm.metrics]
        for ln in em.lines:
            w("  " + ln)
        for k, e in env_exprs.items():
            w(f"  env[{tm.env_names.index(k)}] = (double){e};")
        for i, (k, e) in enumerate(met_exprs):
            w(f"  m[{i}] = (double){e};")
        w("}\n")
    nacc = max(meta["n_acc"], 1)
    meta["n_acc"] = nacc
    # ---- kernels -----------------------------------------------------------------------------
    w(f"constexpr int NACC = {nacc};")
    w('''
template <int MODE, int VAR>
__global__ void __launch_bounds__(kThreads) jxc_step_kernel(const ModelDev md) {
  __shared__ double s_red[(kThreads / 32) * NACC];
  __shared__ double s_tot[NACC];
  __shared__ int s_last;
  int ti = 0;
  for (int i = 1; i < md.n_types; ++i)
    if ((int)blockIdx.x >= md.t[i].block_begin) ti = i;
  const TypeDev& t = md.t[ti];
  const int lb = blockIdx.x - t.block_begin;
  const uint32_t* kp = md.keys + (size_t)md.ctrl->step_in_run * (md.n_types + 1) * 2;
  const Key ck = {kp[2 * ti], kp[2 * ti + 1]};
  const long long time_step = md.ctrl->time_step;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double accd[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) accd[i] = 0.0;
  bool is_minmax[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) is_minmax[i] = false;''')
    for var, tm in enumerate(variants):
        w(f"  if (VAR == {var}) {{")
        w("    switch (ti) {")
        for ti in range(n_types):
            w(f"      case {ti}: jxc_agents_v{var}_t{ti}<MODE>(t, md.env, md.consts, time_step, ck, lb, accd); break;")
        w("    }")
        w("  }")
    w('''  if (lane == 0)
    for (int i = 0; i < NACC; ++i) s_red[warp * NACC + i] = accd[i];
  __syncthreads();''')
    # fold across warps / CTAs needs to know the combine op and owner type of every slot
    for var, tm in enumerate(variants):
        roots = list(tm.env_out.values()) + [v for _, v in tm.metrics]
        reds = used_leaves(roots, "reduce")
        kinds = ", ".join({"sum": "0", "mean": "0", "max": "1", "min": "2"}[r.attr[0]] for r in reds) or "0"
        owners = ", ".join(str(r.attr[1]) for r in reds) or "0"
        w(f"  const int kind_v{var}[NACC] = {{{kinds}}}; const int owner_v{var}[NACC] = {{{owners}}};")
    w("  const int* kind = " + " : ".join([f"VAR == {v} ? kind_v{v}" for v in range(len(variants) - 1)] + [f"kind_v{len(variants) - 1}"]) + ";")
    w("  const int* owner = " + " : ".join([f"VAR == {v} ? owner_v{v}" for v in range(len(variants) - 1)] + [f"owner_v{len(variants) - 1}"]) + ";")
    w('''  // a CTA only contributes to the slots of its own collection; the others get the identity
  if (threadIdx.x < NACC) {
    const int i = threadIdx.x;
    double r = kind[i] == 0 ? 0.0 : (kind[i] == 1 ? -1.0 / 0.0 : 1.0 / 0.0);
    if (owner[i] == ti) {
      r = s_red[i];
      for (int w2 = 1; w2 < kThreads / 32; ++w2) {
        const double v = s_red[w2 * NACC + i];
        r = kind[i] == 0 ? r + v : (kind[i] == 1 ? fmax(r, v) : fmin(r, v));
      }
    }
    md.partials[(size_t)blockIdx.x * NACC + i] = r;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&md.ctrl->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = warp; i < NACC; i += kThreads / 32) {
    double r = kind[i] == 0 ? 0.0 : (kind[i] == 1 ? -1.0 / 0.0 : 1.0 / 0.0);
    for (int b = lane; b < (int)gridDim.x; b += 32) {
      const double v = __ldcg(md.partials + (size_t)b * NACC + i);
      r = kind[i] == 0 ? r + v : (kind[i] == 1 ? fmax(r, v) : fmin(r, v));
    }
    if (kind[i] == 0) r = warp_sum(r);
    else if (kind[i] == 1) r = warp_max(r);
    else r = -warp_max(-r);
    if (lane == 0) s_tot[i] = r;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Ctrl* c = md.ctrl;
    c->ticket = 0;
    const Key uk = {kp[2 * md.n_types], kp[2 * md.n_types + 1]};
    double m[kMaxMetrics];
#pragma unroll
    for (int i = 0; i < kMaxMetrics; ++i) m[i] = 0.0;''')
    for var in range(len(variants)):
        w(f"    if (VAR == {var}) jxc_tail_v{var}<MODE>(md, s_tot, uk, m);")
    w('''    const long long tn = c->time_step + 1;
    if ((tn % md.collect_interval) == 0) {
      double* row = md.metrics + (size_t)c->n_recorded * kMaxMetrics;
#pragma unroll
      for (int i = 0; i < kMaxMetrics; ++i) row[i] = m[i];
      md.record_steps[c->n_recorded] = (int)tn;
      c->n_recorded += 1;
    }
    c->time_step = tn;
    c->step_in_run += 1;
  }
}
''')
    # ---- init (AgentCollection.init, agent.py:92-130) -------------------------------------------
    for ti, t in enumerate(tm0.types):
        w(f"template <int MODE> __device__ __forceinline__ void jxc_init_one_t{ti}(const TypeDev& t, Key key, "
          f"const double* __restrict__ cst, long long i) {{")
        w("  const Key ak = split_child<MODE>(key, (unsigned long long)(t.goff + i), (unsigned long long)t.gn);")

        def leaf_init(x, em):
            raise TraceError("init_state can only use its key and Python constants")
        em = Emitter(leaf_init, "ak", pool)
        vals = [(fi, dt, wd, c, em.cast(t["init"][fname][c], dt)) for fi, (fname, dt, wd) in enumerate(t["fields"]) for c in range(wd)]
        for ln in em.lines:
            w("  " + ln)
        for fi, dt, wd, c, e in vals:
            ct = {F32: "float", I32: "int", BOOL: "unsigned char"}[dt]
            w(f"  (({ct}*)t.f[{fi}])[i * {wd} + {c}] = {e}{' ? 1 : 0' if dt == BOOL else ''};")
        w("}")
        w(f"template <int MODE> __global__ void __launch_bounds__(kThreads) jxc_init_kernel_t{ti}(const TypeDev t, Key key, "
          f"const double* __restrict__ cst) {{")
        w("  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < t.n; i += (long long)gridDim.x * blockDim.x)")
        w(f"    jxc_init_one_t{ti}<MODE>(t, key, cst, i);")
        w("}\n")
    # ---- launchers ------------------------------------------------------------------------------
    w('extern "C" int jxc_n_acc() { return NACC; }')
    w(f'extern "C" int jxc_n_variants() {{ return {len(variants)}; }}')
    w('extern "C" int jxc_launch_init(const ModelDev* md, int type, unsigned int k0, unsigned int k1, int rng_mode, int blocks, cudaStream_t s) {')
    w("  const Key key{k0, k1};\n  switch (type) {")
    for ti in range(n_types):
        w(f"    case {ti}: if (rng_mode == 1) jxc_init_kernel_t{ti}<1><<<blocks, kThreads, 0, s>>>(md->t[{ti}], key, md->consts); "
          f"else jxc_init_kernel_t{ti}<0><<<blocks, kThreads, 0, s>>>(md->t[{ti}], key, md->consts); break;")
    w("    default: return -1;\n  }\n  return (int)cudaGetLastError();\n}")
    w('extern "C" int jxc_launch_step(const ModelDev* md, int rng_mode, int variant, cudaStream_t s) {')
    for var in range(len(variants)):
        w(f"  if (variant == {var}) {{ if (rng_mode == 1) jxc_step_kernel<1, {var}><<<md->grid_blocks, kThreads, 0, s>>>(*md); "
          f"else jxc_step_kernel<0, {var}><<<md->grid_blocks, kThreads, 0, s>>>(*md); }}")
    w("  return (int)cudaGetLastError();\n}")
    # ---- replica-parallel ensemble of the traced model (analysis.py:113-157, :434-476) ----------------
    kinds_last = ", ".join({"sum": "0", "mean": "0", "max": "1", "min": "2"}[r.attr[0]] for r in
                           used_leaves(list(variants[-1].env_out.values()) + [v for _, v in variants[-1].metrics], "reduce")) or "0"
    kinds_first = ", ".join({"sum": "0", "mean": "0", "max": "1", "min": "2"}[r.attr[0]] for r in
                            used_leaves(list(variants[0].env_out.values()) + [v for _, v in variants[0].metrics], "reduce")) or "0"
    field_sizes = [[{F32: 4, I32: 4, BOOL: 1}[dt] * wd for _, dt, wd in t["fields"]] for t in tm0.types]      # bytes per agent
    w(f"constexpr int kEnsTypes = {n_types};")
    w("constexpr int kEnsThreads2 = 1024;")
    w("struct JxcEns { int R, steps, n_consts, has_env_fn, use_smem; const double* consts; const double* env0; const unsigned int* seeds; "
      "double* out; unsigned char* scratch; size_t state_bytes; long long n[JXB_MAX_TYPES]; size_t foff[JXB_MAX_TYPES][kMaxFields]; };")
    w("""
// One CTA owns one replica at a time and runs its whole life in shared memory (or an L2-resident slot):
// initialize from PRNGKey(seed), then `steps` x (traced agent updates -> reduction -> traced env/metrics
// tail), with the key schedule of model.py:129-130,156,164,183 derived in the kernel.
template <int MODE>
__global__ void __launch_bounds__(kEnsThreads2) jxc_ensemble_kernel(const JxcEns e) {
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ ModelDev md;
  __shared__ Ctrl ctrl;
  __shared__ double env[kMaxEnv];
  __shared__ double s_red[(kEnsThreads2 / 32) * NACC];
  __shared__ double s_tot[NACC];
  __shared__ double metrics[kMaxMetrics];
  __shared__ Key s_keys[JXB_MAX_TYPES + 2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* base = e.use_smem ? dsm : e.scratch + (size_t)blockIdx.x * e.state_bytes;
  for (int r = blockIdx.x; r < e.R; r += gridDim.x) {
    const double* cst = e.consts + (size_t)r * e.n_consts;
    if (tid == 0) {
      md.n_types = kEnsTypes; md.collect_interval = 1; md.world_size = 1; md.env = env; md.ctrl = &ctrl; md.consts = cst;
      for (int i = 0; i < kEnsTypes; ++i) {
        md.t[i].n = e.n[i]; md.t[i].goff = 0; md.t[i].gn = e.n[i]; md.t[i].block_begin = 0; md.t[i].block_count = 1;
        for (int f = 0; f < kMaxFields; ++f) md.t[i].f[f] = base + e.foff[i][f];
      }
      for (int k = 0; k < kMaxEnv; ++k) env[k] = e.env0[(size_t)r * kMaxEnv + k];
      for (int k = 0; k < kMaxMetrics; ++k) metrics[k] = 0.0;
      ctrl.time_step = 0;
      const Key root{0u, e.seeds[r]};
      s_keys[JXB_MAX_TYPES + 1] = split_child<MODE>(root, 0ull, (unsigned long long)(kEnsTypes + 1));     // model _rng
      for (int i = 0; i < kEnsTypes; ++i) s_keys[i] = split_child<MODE>(root, (unsigned long long)(i + 1), (unsigned long long)(kEnsTypes + 1));
    }
    __syncthreads();""")
    for ti in range(n_types):
        w(f"    for (long long i = tid; i < md.t[{ti}].n; i += blockDim.x) jxc_init_one_t{ti}<MODE>(md.t[{ti}], s_keys[{ti}], cst, i);")
    w("""    __syncthreads();
    for (int step = 0; step < e.steps; ++step) {
      if (tid == 0) {                       // model.py:156,164,183
        Key rng = s_keys[JXB_MAX_TYPES + 1];
        Key step_key = split_child<MODE>(rng, 1ull, 2ull);
        s_keys[JXB_MAX_TYPES + 1] = split_child<MODE>(rng, 0ull, 2ull);
        for (int i = 0; i < kEnsTypes; ++i) {
          s_keys[i] = split_child<MODE>(step_key, 1ull, 2ull);
          step_key = split_child<MODE>(step_key, 0ull, 2ull);
        }
        s_keys[JXB_MAX_TYPES] = e.has_env_fn ? split_child<MODE>(step_key, 1ull, 2ull) : Key{0u, 0u};
      }
      __syncthreads();
      double accd[NACC];
#pragma unroll
      for (int i = 0; i < NACC; ++i) accd[i] = 0.0;
      const long long time_step = ctrl.time_step;
      const bool first = step == 0;""")
    nv = len(variants)
    for ti in range(n_types):
        if nv > 1:
            w(f"      if (first) jxc_agents_v0_t{ti}<MODE>(md.t[{ti}], env, cst, time_step, s_keys[{ti}], 0, accd);")
            w(f"      else jxc_agents_v{nv - 1}_t{ti}<MODE>(md.t[{ti}], env, cst, time_step, s_keys[{ti}], 0, accd);")
        else:
            w(f"      jxc_agents_v0_t{ti}<MODE>(md.t[{ti}], env, cst, time_step, s_keys[{ti}], 0, accd);")
    w(f"      const int kind0[NACC] = {{{kinds_first}}}; const int kind1[NACC] = {{{kinds_last}}};")
    w("""      const int* kind = first ? kind0 : kind1;
      if (lane == 0)
        for (int i = 0; i < NACC; ++i) s_red[warp * NACC + i] = accd[i];
      __syncthreads();
      if (tid < NACC) {
        double rr = s_red[tid];
        for (int w2 = 1; w2 < kEnsThreads2 / 32; ++w2) {
          const double v = s_red[w2 * NACC + tid];
          rr = kind[tid] == 0 ? rr + v : (kind[tid] == 1 ? fmax(rr, v) : fmin(rr, v));
        }
        s_tot[tid] = rr;
      }
      __syncthreads();
      if (tid == 0) {""")
    if nv > 1:
        w(f"        if (first) jxc_tail_v0<MODE>(md, s_tot, s_keys[JXB_MAX_TYPES], metrics); else jxc_tail_v{nv - 1}<MODE>(md, s_tot, s_keys[JXB_MAX_TYPES], metrics);")
    else:
        w("        jxc_tail_v0<MODE>(md, s_tot, s_keys[JXB_MAX_TYPES], metrics);")
    w("""        ctrl.time_step += 1;
      }
      __syncthreads();
    }
    if (tid < kMaxMetrics) e.out[(size_t)r * kMaxMetrics + tid] = metrics[tid];
    __syncthreads();
  }
}
""")
    sizes_init = "{" + ", ".join("{" + ", ".join(str(x) for x in fs) + "}" for fs in field_sizes) + "}"
    nf_init = "{" + ", ".join(str(len(fs)) for fs in field_sizes) + "}"
    w(f"static const int kFieldSize[kEnsTypes][kMaxFields] = {sizes_init};")
    w(f"static const int kNumFields[kEnsTypes] = {nf_init};")
    w("""
// Self-contained ensemble entry point: R replicas, replica r uses the constant table consts[r][n_consts], the
// env values env0[r][kMaxEnv] and PRNGKey(seeds[r]); last_metrics[r][kMaxMetrics] receives results[m][-1].
extern "C" int jxc_ensemble_run(int device, int R, int steps, const long long* n_agents, const double* consts, int n_consts,
                                const double* env0, const unsigned int* seeds, int has_env_fn, int rng_mode,
                                double* last_metrics, double* device_seconds) {
  if (cudaSetDevice(device) != cudaSuccess) return -1;
  JxcEns e;
  memset(&e, 0, sizeof(e));
  e.R = R; e.steps = steps; e.n_consts = n_consts; e.has_env_fn = has_env_fn;
  size_t off = 0;
  for (int i = 0; i < kEnsTypes; ++i) {
    e.n[i] = n_agents[i];
    for (int f = 0; f < kNumFields[i]; ++f) {
      e.foff[i][f] = off;
      off += (((size_t)n_agents[i] + 8) * kFieldSize[i][f] + 15) / 16 * 16;
    }
  }
  e.state_bytes = off;
  e.use_smem = off <= 200 * 1024;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -2;
  int grid = e.use_smem ? (R < prop.multiProcessorCount ? R : prop.multiProcessorCount)
                        : (R < 2 * prop.multiProcessorCount ? R : 2 * prop.multiProcessorCount);
  double *d_c = nullptr, *d_e = nullptr, *d_o = nullptr; unsigned int* d_s = nullptr; unsigned char* d_scr = nullptr;
  cudaEvent_t e0, e1;
  int rc = 0;
#define JCK(x) do { if ((x) != cudaSuccess) { rc = -3; goto done; } } while (0)
  JCK(cudaMalloc(&d_c, (size_t)R * (n_consts > 0 ? n_consts : 1) * sizeof(double)));
  JCK(cudaMalloc(&d_e, (size_t)R * kMaxEnv * sizeof(double)));
  JCK(cudaMalloc(&d_o, (size_t)R * kMaxMetrics * sizeof(double)));
  JCK(cudaMalloc(&d_s, (size_t)R * sizeof(unsigned int)));
  if (!e.use_smem) JCK(cudaMalloc(&d_scr, (size_t)grid * off));
  if (n_consts) JCK(cudaMemcpy(d_c, consts, (size_t)R * n_consts * sizeof(double), cudaMemcpyHostToDevice));
  JCK(cudaMemcpy(d_e, env0, (size_t)R * kMaxEnv * sizeof(double), cudaMemcpyHostToDevice));
  JCK(cudaMemcpy(d_s, seeds, (size_t)R * sizeof(unsigned int), cudaMemcpyHostToDevice));
  e.consts = d_c; e.env0 = d_e; e.seeds = d_s; e.out = d_o; e.scratch = d_scr;
  {
    const size_t dyn = e.use_smem ? off : 0;
    if (dyn > 48 * 1024) {
      JCK(cudaFuncSetAttribute(jxc_ensemble_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
      JCK(cudaFuncSetAttribute(jxc_ensemble_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    }
    JCK(cudaEventCreate(&e0)); JCK(cudaEventCreate(&e1));
    JCK(cudaEventRecord(e0, 0));
    if (rng_mode == 1) jxc_ensemble_kernel<1><<<grid, kEnsThreads2, dyn>>>(e);
    else jxc_ensemble_kernel<0><<<grid, kEnsThreads2, dyn>>>(e);
    JCK(cudaGetLastError());
    JCK(cudaEventRecord(e1, 0));
    JCK(cudaMemcpy(last_metrics, d_o, (size_t)R * kMaxMetrics * sizeof(double), cudaMemcpyDeviceToHost));
    float ms = 0.f;
    JCK(cudaEventElapsedTime(&ms, e0, e1));
    if (device_seconds) *device_seconds = ms * 1e-3;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
done:
#undef JCK
  cudaFree(d_c); cudaFree(d_e); cudaFree(d_o); cudaFree(d_s); cudaFree(d_scr);
  return rc;
}
""")
    meta["consts"] = pool
    return "\n".join(out) + "\n", meta


def _env_load(x: Tr, arr: str) -> str:
    slot, dt = x.attr, x.dtype
    if dt == WF:
        return f"{arr}[{slot}]"
    if dt == F32:
        return f"(float){arr}[{slot}]"
    if dt in (I32, WI):
        return f"(int){arr}[{slot}]"
    return f"({arr}[{slot}] != 0.0)"
