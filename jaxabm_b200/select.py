"""``AgentCollection.filter`` conditions -> device predicate programs (``csrc/record.cuh``).

The reference evaluates ``condition(states)`` eagerly with ``jax.numpy`` and gathers ``v[mask]`` column by column
(``jaxabm/agent.py:213-243``).  Here the condition is run ONCE on symbolic columns (the tracer's :class:`Tr`
values, JAX's x64-disabled promotion rules) and the recorded expression becomes a postfix program that the
compaction kernel evaluates per agent -- no column leaves HBM.  Conditions the tracer cannot follow (NumPy ufuncs
called on the columns, Python control flow) raise :class:`TraceError`; the caller then evaluates the mask on
lazily downloaded columns and still compacts on the device.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np

from . import trace as T

# opcodes: keep in step with the JP_* enum of csrc/record.cuh
_OPS = ("LOAD_F32 LOAD_I32 LOAD_U8 CONST_F32 CONST_I32 ADD_F SUB_F MUL_F DIV_F MIN_F MAX_F ADD_I SUB_I MUL_I MIN_I MAX_I "
        "NEG_F NEG_I ABS_F ABS_I I2F F2I B2I B2F I2B F2B LT_F LE_F GT_F GE_F EQ_F NE_F LT_I LE_I GT_I GE_I EQ_I NE_I "
        "AND OR XOR NOT SELECT SQRT_F EXP_F LOG_F").split()
OP = {name: i + 1 for i, name in enumerate(_OPS)}
MAX_INS = 96


def _kind(dt: str) -> str:
    return {"f32": "F", "wf64": "F", "i32": "I", "wi32": "I", "bool": "B"}[dt]


class _WideColumn:
    """A ``(N, w)`` state column inside a traced condition: ``col[:, k]`` is component k of every agent."""

    def __init__(self, field: int, dtype: str, width: int):
        self.field, self.dtype, self.width = field, dtype, width

    def __getitem__(self, idx):
        if isinstance(idx, tuple) and len(idx) == 2 and idx[0] in (slice(None), Ellipsis) and isinstance(idx[1], (int, np.integer)):
            k = int(idx[1])
            if k < 0:
                k += self.width
            if 0 <= k < self.width:
                return T.Tr("field", (), self.dtype, "a", (0, self.field, k))
        raise T.TraceError("only col[:, k] is traced on a vector-valued state column")


def compile_predicate(condition: Callable, fields: Sequence[Tuple[str, type, int]]) -> List[Tuple[int, int, int, float]]:
    """Trace ``condition(states)`` -> [(op, a, b, f)] in postfix order.  Raises :class:`TraceError`."""
    dts = {np.float32: T.F32, np.int32: T.I32, np.bool_: T.BOOL}
    cols = {}
    for fi, (name, dt, w) in enumerate(fields):
        tdt = dts[np.dtype(dt).type]
        cols[name] = T.Tr("field", (), tdt, "a", (0, fi, 0)) if w == 1 else _WideColumn(fi, tdt, w)
    try:
        out = condition(cols)
    except T.TraceError:
        raise
    except Exception as e:          # NumPy refusing a symbolic operand, Python control flow, ...
        raise T.TraceError(f"condition is not traceable: {type(e).__name__}: {e}") from e
    if not isinstance(out, T.Tr) or out.dtype != T.BOOL:
        raise T.TraceError("condition must return a boolean expression of the state columns")
    prog: List[Tuple[int, int, int, float]] = []

    def emit(op: str, a: int = 0, b: int = 0, f: float = 0.0):
        prog.append((OP[op], int(a), int(b), float(f)))
        if len(prog) > MAX_INS:
            raise T.TraceError(f"condition compiles to more than {MAX_INS} instructions")

    def cast(x: T.Tr, to: str):
        if x.op == "const" and x.dtype != T.BOOL:          # weak or typed scalar: materialise directly in the target kind
            if to == "F":
                return emit("CONST_F32", f=float(np.float32(x.attr)))
            if to == "I":
                return emit("CONST_I32", a=int(x.attr))
            return emit("CONST_I32", a=int(bool(x.attr)))
        walk(x)
        k = _kind(x.dtype)
        if k != to:
            emit({"IF": "I2F", "FI": "F2I", "BI": "B2I", "BF": "B2F", "IB": "I2B", "FB": "F2B"}[k + to])

    def walk(x: T.Tr):
        op, k = x.op, _kind(x.dtype)
        if op == "const":
            if k == "F":
                emit("CONST_F32", f=float(np.float32(x.attr)))
            else:
                emit("CONST_I32", a=int(x.attr))
        elif op == "field":
            emit({"F": "LOAD_F32", "I": "LOAD_I32", "B": "LOAD_U8"}[k], a=x.attr[1], b=x.attr[2] if len(x.attr) > 2 else 0)
        elif op in ("add", "sub", "mul", "div", "min", "max"):
            if k == "B" or (op == "div" and k != "F"):
                raise T.TraceError(f"{op} on {x.dtype} is not traced in a filter condition")
            cast(x.args[0], k), cast(x.args[1], k)
            emit(f"{op.upper()}_{k}")
        elif op in ("neg", "abs"):
            cast(x.args[0], k)
            emit(f"{op.upper()}_{k}")
        elif op in ("lt", "le", "gt", "ge", "eq", "ne"):
            ck = _kind(x.attr)
            ck = "I" if ck == "B" else ck
            cast(x.args[0], ck), cast(x.args[1], ck)
            emit(f"{op.upper()}_{ck}")
        elif op in ("and", "or", "xor"):
            walk(x.args[0]), walk(x.args[1])
            emit(op.upper())
        elif op == "not":
            walk(x.args[0])
            emit("NOT")
        elif op == "where":
            walk(x.args[0]), cast(x.args[1], k), cast(x.args[2], k)
            emit("SELECT")
        elif op == "cast":
            cast(x.args[0], k)
        elif op in ("sqrt", "exp", "log"):
            cast(x.args[0], "F")
            emit(f"{op.upper()}_F")
        else:
            raise T.TraceError(f"{op!r} is not traced in a filter condition")

    walk(out)
    return prog
