"""ctypes binding of the CPU oracle (oracle/libfq_oracle.so).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by anything under fuse_query_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Any, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfq_oracle.so")

# tags follow datavalues/data_value.rs:19-35
NULL, BOOL, I8, I16, I32, I64, U8, U16, U32, U64, F32, F64, UTF8, STRUCT = range(14)
DTYPE_NAMES = ["Null", "Boolean", "Int8", "Int16", "Int32", "Int64", "UInt8", "UInt16",
               "UInt32", "UInt64", "Float32", "Float64", "Utf8", "Struct"]
NP_OF = {BOOL: np.uint8, I8: np.int8, I16: np.int16, I32: np.int32, I64: np.int64, U8: np.uint8,
         U16: np.uint16, U32: np.uint32, U64: np.uint64, F32: np.float32, F64: np.float64}
TAG_OF_NP = {np.dtype(v): k for k, v in NP_OF.items() if k != BOOL}
TAG_OF_NP[np.dtype(np.bool_)] = BOOL

AGG = {"min": 0, "max": 1, "sum": 2, "count": 3}
CMP = {"=": 0, "<": 1, "<=": 2, ">": 3, ">=": 4}
ARITH = {"+": 0, "-": 1, "*": 2, "/": 3}
LOGIC = {"and": 0, "or": 1}

ERRLEN = 512
MAX_COLS = 16


class _ValueU(C.Union):
    _fields_ = [("i", C.c_int64), ("u", C.c_uint64), ("f", C.c_double)]


class CValue(C.Structure):
    pass


CValue._fields_ = [("tag", C.c_int32), ("some", C.c_int32), ("v", _ValueU), ("s", C.c_void_p),
                   ("items", C.POINTER(CValue)), ("n_items", C.c_int32), ("_pad", C.c_int32)]


class CArray(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("owned", C.c_int32), ("len", C.c_int64), ("data", C.c_void_p),
                ("valid", C.c_void_p)]


class CColumnar(C.Structure):
    _fields_ = [("is_scalar", C.c_int32), ("_pad", C.c_int32), ("scalar", CValue), ("array", CArray)]


class CBlock(C.Structure):
    _fields_ = [("n_cols", C.c_int32), ("_pad", C.c_int32), ("names", C.c_char_p * MAX_COLS),
                ("cols", CArray * MAX_COLS)]


class CQuery(C.Structure):
    _fields_ = [("total", C.c_uint64), ("table", C.POINTER(CBlock)), ("block_size", C.c_uint64),
                ("tail_quirk", C.c_int32), ("worker_threads", C.c_int32), ("use_threads", C.c_int32),
                ("fused", C.c_int32), ("predicate", C.c_char_p), ("n_exprs", C.c_int32),
                ("is_aggregate", C.c_int32), ("exprs", C.POINTER(C.c_char_p)), ("limit", C.c_int64)]


class CResult(C.Structure):
    _fields_ = [("block", CBlock), ("n_rows", C.c_int64), ("n_blocks_out", C.c_int64),
                ("rows_scanned", C.c_int64), ("partial_states_json", C.c_void_p), ("seconds", C.c_double)]


class OracleError(Exception):
    """Carries the reference's Display text of FuseQueryError (error.rs:10-20)."""


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, "fq_oracle.c"), os.path.join(_HERE, "fq_oracle.h")]
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libfq_oracle.so"], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_generate_parts.argtypes = [C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_generate_parts.restype = C.c_int32
        L.orc_block_ranges.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_uint64), C.c_int64]
        L.orc_block_ranges.restype = C.c_int64
        for name in ("orc_array_arithmetic", "orc_array_comparison", "orc_array_logic"):
            f = getattr(L, name)
            f.argtypes = [C.c_int32, C.POINTER(CColumnar), C.POINTER(CColumnar), C.POINTER(CArray), C.c_char_p]
            f.restype = C.c_int32
        L.orc_array_aggregate.argtypes = [C.c_int32, C.POINTER(CArray), C.POINTER(CValue), C.c_char_p]
        L.orc_array_aggregate.restype = C.c_int32
        for name in ("orc_value_arithmetic", "orc_value_aggregate"):
            f = getattr(L, name)
            f.argtypes = [C.c_int32, C.POINTER(CValue), C.POINTER(CValue), C.POINTER(CValue), C.c_char_p]
            f.restype = C.c_int32
        L.orc_numerical_coercion.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_char_p]
        L.orc_numerical_coercion.restype = C.c_int32
        L.orc_value_to_json.argtypes = [C.POINTER(CValue)]
        L.orc_value_to_json.restype = C.c_void_p
        L.orc_value_from_json.argtypes = [C.c_char_p, C.POINTER(CValue), C.c_char_p]
        L.orc_value_from_json.restype = C.c_int32
        L.orc_value_display.argtypes = [C.POINTER(CValue)]
        L.orc_value_display.restype = C.c_void_p
        L.orc_value_free.argtypes = [C.POINTER(CValue)]
        L.orc_array_free.argtypes = [C.POINTER(CArray)]
        L.orc_columnar_free.argtypes = [C.POINTER(CColumnar)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_fn_parse.argtypes = [C.c_char_p, C.c_char_p]
        L.orc_fn_parse.restype = C.c_void_p
        L.orc_fn_clone.argtypes = [C.c_void_p]
        L.orc_fn_clone.restype = C.c_void_p
        L.orc_fn_free.argtypes = [C.c_void_p]
        L.orc_fn_display.argtypes = [C.c_void_p]
        L.orc_fn_display.restype = C.c_void_p
        L.orc_plan_display.argtypes = [C.c_char_p, C.c_char_p]
        L.orc_plan_display.restype = C.c_void_p
        L.orc_fn_is_aggregate.argtypes = [C.c_void_p]
        L.orc_fn_is_aggregate.restype = C.c_int32
        L.orc_fn_set_depth.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_fn_return_type.argtypes = [C.c_void_p, C.POINTER(CBlock), C.POINTER(C.c_int32), C.c_char_p]
        L.orc_fn_return_type.restype = C.c_int32
        L.orc_fn_eval.argtypes = [C.c_void_p, C.POINTER(CBlock), C.POINTER(CColumnar), C.c_char_p]
        L.orc_fn_eval.restype = C.c_int32
        L.orc_fn_accumulate.argtypes = [C.c_void_p, C.POINTER(CBlock), C.c_char_p]
        L.orc_fn_accumulate.restype = C.c_int32
        L.orc_fn_accumulate_result.argtypes = [C.c_void_p, C.POINTER(CValue), C.c_char_p]
        L.orc_fn_accumulate_result.restype = C.c_int32
        L.orc_fn_merge_state.argtypes = [C.c_void_p, C.POINTER(CValue), C.c_char_p]
        L.orc_fn_merge_state.restype = C.c_int32
        L.orc_fn_merge_result.argtypes = [C.c_void_p, C.POINTER(CValue), C.c_char_p]
        L.orc_fn_merge_result.restype = C.c_int32
        L.orc_query_run.argtypes = [C.POINTER(CQuery), C.POINTER(CResult), C.c_char_p]
        L.orc_query_run.restype = C.c_int32
        L.orc_result_free.argtypes = [C.POINTER(CResult)]
        _lib = L
    return _lib


# ---------------------------------------------------------------------------------------------
# Python-side values
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Value:
    """DataValue.  tag == NULL is DataValue::Null; value None is Type(None)."""
    tag: int
    value: Any = None

    def __repr__(self):
        if self.tag == NULL:
            return "Null"
        return f"{DTYPE_NAMES[self.tag]}({self.value!r})"


def Null() -> Value:
    return Value(NULL)


def _take_str(p) -> str:
    s = C.string_at(p).decode()
    lib().orc_free(p)
    return s


def _to_cvalue(v: Value, keep: list) -> CValue:
    cv = CValue()
    cv.tag = v.tag
    if v.tag == NULL:
        return cv
    if v.tag == STRUCT:
        items = (CValue * max(1, len(v.value)))()
        for i, it in enumerate(v.value):
            items[i] = _to_cvalue(it, keep)
        keep.append(items)
        cv.items = C.cast(items, C.POINTER(CValue))
        cv.n_items = len(v.value)
        return cv
    if v.value is None:
        return cv
    cv.some = 1
    if v.tag == UTF8:
        b = C.create_string_buffer(v.value.encode())
        keep.append(b)
        cv.s = C.cast(b, C.c_void_p)
    elif v.tag in (F32, F64):
        cv.v.f = float(np.float32(v.value)) if v.tag == F32 else float(v.value)
    elif v.tag in (U8, U16, U32, U64):
        cv.v.u = int(v.value)
    else:
        cv.v.i = int(v.value)
    return cv


def _from_cvalue(cv: CValue) -> Value:
    if cv.tag == NULL:
        return Value(NULL)
    if cv.tag == STRUCT:
        return Value(STRUCT, tuple(_from_cvalue(cv.items[i]) for i in range(cv.n_items)))
    if not cv.some:
        return Value(cv.tag, None)
    if cv.tag == UTF8:
        return Value(UTF8, C.string_at(cv.s).decode())
    if cv.tag in (F32, F64):
        return Value(cv.tag, float(cv.v.f))
    if cv.tag in (U8, U16, U32, U64):
        return Value(cv.tag, int(cv.v.u))
    if cv.tag == BOOL:
        return Value(BOOL, bool(cv.v.i))
    return Value(cv.tag, int(cv.v.i))


@dataclass
class Array:
    """Arrow-style array on the host: numpy values (+ optional validity), or list[str|None] for Utf8."""
    dtype: int
    values: Any
    valid: Optional[np.ndarray] = None

    def __len__(self):
        return len(self.values)

    def to_list(self) -> list:
        if self.dtype == UTF8:
            return list(self.values)
        vals = self.values.tolist()
        if self.dtype == BOOL:
            vals = [bool(x) for x in vals]
        if self.valid is None:
            return vals
        return [v if ok else None for v, ok in zip(vals, self.valid.tolist())]


def array(dtype: int, values: Sequence, valid: Optional[Sequence] = None) -> Array:
    if dtype == UTF8:
        return Array(UTF8, list(values))
    a = np.ascontiguousarray(np.asarray(values, dtype=NP_OF[dtype]))
    v = None if valid is None else np.ascontiguousarray(np.asarray(valid, dtype=np.uint8))
    return Array(dtype, a, v)


def from_numpy(a: np.ndarray) -> Array:
    a = np.ascontiguousarray(a)
    tag = TAG_OF_NP[a.dtype]
    if tag == BOOL:
        a = a.astype(np.uint8)
    return Array(tag, a)


def _to_carray(a: Array, keep: list) -> CArray:
    ca = CArray()
    ca.dtype = a.dtype
    ca.owned = 0
    ca.len = len(a.values)
    if a.dtype == UTF8:
        bufs = [None if s is None else C.create_string_buffer(s.encode()) for s in a.values]
        ptrs = (C.c_void_p * max(1, len(bufs)))(*[None if b is None else C.cast(b, C.c_void_p) for b in bufs])
        keep.extend([bufs, ptrs])
        ca.data = C.cast(ptrs, C.c_void_p)
    else:
        keep.append(a.values)
        ca.data = a.values.ctypes.data if len(a.values) else None
    if a.valid is not None:
        keep.append(a.valid)
        ca.valid = a.valid.ctypes.data
    return ca


def _from_carray(ca: CArray) -> Array:
    n = ca.len
    if ca.dtype == NULL:
        return Array(NULL, np.zeros(n, np.uint8), np.zeros(n, np.uint8))
    valid = None
    if ca.valid:
        valid = np.ctypeslib.as_array(C.cast(ca.valid, C.POINTER(C.c_uint8)), shape=(n,)).copy() if n else np.zeros(0, np.uint8)
    if ca.dtype == UTF8:
        ptrs = C.cast(ca.data, C.POINTER(C.c_void_p))
        vals = [None if not ptrs[i] else C.string_at(ptrs[i]).decode() for i in range(n)]
        if valid is not None:
            vals = [v if ok else None for v, ok in zip(vals, valid)]
        return Array(UTF8, vals)
    npdt = NP_OF[ca.dtype]
    if n == 0:
        return Array(ca.dtype, np.zeros(0, npdt), valid)
    ct = np.ctypeslib.as_ctypes_type(npdt)
    vals = np.ctypeslib.as_array(C.cast(ca.data, C.POINTER(ct)), shape=(n,)).copy()
    return Array(ca.dtype, vals, valid)


def _columnar(x, keep: list) -> CColumnar:
    cc = CColumnar()
    if isinstance(x, Value):
        cc.is_scalar = 1
        cc.scalar = _to_cvalue(x, keep)
    else:
        cc.is_scalar = 0
        cc.array = _to_carray(x, keep)
    return cc


def _err():
    return C.create_string_buffer(ERRLEN)


def _raise(e):
    raise OracleError(e.value.decode())


def _block(cols: dict, keep: list) -> CBlock:
    b = CBlock()
    b.n_cols = len(cols)
    for i, (name, arr) in enumerate(cols.items()):
        b.names[i] = name.encode()
        b.cols[i] = _to_carray(arr, keep)
    return b


# ---------------------------------------------------------------------------------------------
# datavalues
# ---------------------------------------------------------------------------------------------
def generate_parts(total: int):
    b = (C.c_uint64 * 8)()
    e = (C.c_uint64 * 8)()
    n = lib().orc_generate_parts(total, b, e)
    return [(int(b[i]), int(e[i])) for i in range(n)]


def block_ranges(begin: int, end: int, block_size: int = 10000, tail_quirk: bool = True):
    n = lib().orc_block_ranges(begin, end, block_size, int(tail_quirk), None, None, 0)
    b = (C.c_uint64 * n)()
    e = (C.c_uint64 * n)()
    lib().orc_block_ranges(begin, end, block_size, int(tail_quirk), b, e, n)
    return [(int(b[i]), int(e[i])) for i in range(n)]


def _binary_array_op(fn, op: int, left, right) -> Array:
    keep: list = []
    l, r = _columnar(left, keep), _columnar(right, keep)
    out, e = CArray(), _err()
    if fn(op, C.byref(l), C.byref(r), C.byref(out), e):
        _raise(e)
    res = _from_carray(out)
    lib().orc_array_free(C.byref(out))
    return res


def array_arithmetic(op: str, left, right) -> Array:
    return _binary_array_op(lib().orc_array_arithmetic, ARITH[op], left, right)


def array_comparison(op: str, left, right) -> Array:
    return _binary_array_op(lib().orc_array_comparison, CMP[op], left, right)


def array_logic(op: str, left, right) -> Array:
    return _binary_array_op(lib().orc_array_logic, LOGIC[op], left, right)


def array_aggregate(op: str, a: Array) -> Value:
    keep: list = []
    ca = _to_carray(a, keep)
    out, e = CValue(), _err()
    if lib().orc_array_aggregate(AGG[op], C.byref(ca), C.byref(out), e):
        _raise(e)
    v = _from_cvalue(out)
    lib().orc_value_free(C.byref(out))
    return v


def _binary_value_op(fn, op: int, l: Value, r: Value) -> Value:
    keep: list = []
    cl, cr = _to_cvalue(l, keep), _to_cvalue(r, keep)
    out, e = CValue(), _err()
    if fn(op, C.byref(cl), C.byref(cr), C.byref(out), e):
        _raise(e)
    v = _from_cvalue(out)
    lib().orc_value_free(C.byref(out))
    return v


def value_arithmetic(op: str, l: Value, r: Value) -> Value:
    return _binary_value_op(lib().orc_value_arithmetic, ARITH[op], l, r)


def value_aggregate(op: str, l: Value, r: Value) -> Value:
    return _binary_value_op(lib().orc_value_aggregate, AGG[op], l, r)


def numerical_coercion(op: str, l: int, r: int) -> int:
    out, e = C.c_int32(), _err()
    if lib().orc_numerical_coercion(op.encode(), l, r, C.byref(out), e):
        _raise(e)
    return out.value


def value_to_json(v: Value) -> str:
    keep: list = []
    cv = _to_cvalue(v, keep)
    return _take_str(lib().orc_value_to_json(C.byref(cv)))


def value_from_json(s: str) -> Value:
    out, e = CValue(), _err()
    if lib().orc_value_from_json(s.encode(), C.byref(out), e):
        _raise(e)
    v = _from_cvalue(out)
    lib().orc_value_free(C.byref(out))
    return v


def value_display(v: Value) -> str:
    keep: list = []
    cv = _to_cvalue(v, keep)
    return _take_str(lib().orc_value_display(C.byref(cv)))


def plan_display(sexpr: str) -> str:
    e = _err()
    p = lib().orc_plan_display(sexpr.encode(), e)
    if not p:
        _raise(e)
    return _take_str(p)


# ---------------------------------------------------------------------------------------------
# enum Function
# ---------------------------------------------------------------------------------------------
class Function:
    def __init__(self, sexpr: Optional[str] = None, _h=None):
        if _h is None:
            e = _err()
            _h = lib().orc_fn_parse(sexpr.encode(), e)
            if not _h:
                _raise(e)
        self._h = _h

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_fn_free(self._h)
            self._h = None

    def clone(self) -> "Function":
        return Function(_h=lib().orc_fn_clone(self._h))

    def display(self) -> str:
        return _take_str(lib().orc_fn_display(self._h))

    def is_aggregate(self) -> bool:
        return bool(lib().orc_fn_is_aggregate(self._h))

    def set_depth(self, d: int):
        lib().orc_fn_set_depth(self._h, d)

    def return_type(self, cols: dict) -> int:
        keep: list = []
        b = _block(cols, keep)
        out, e = C.c_int32(), _err()
        if lib().orc_fn_return_type(self._h, C.byref(b), C.byref(out), e):
            _raise(e)
        return out.value

    def eval(self, cols: dict):
        keep: list = []
        b = _block(cols, keep)
        out, e = CColumnar(), _err()
        if lib().orc_fn_eval(self._h, C.byref(b), C.byref(out), e):
            _raise(e)
        res = _from_cvalue(out.scalar) if out.is_scalar else _from_carray(out.array)
        lib().orc_columnar_free(C.byref(out))
        return res

    def accumulate(self, cols: dict):
        keep: list = []
        b = _block(cols, keep)
        e = _err()
        if lib().orc_fn_accumulate(self._h, C.byref(b), e):
            _raise(e)

    def accumulate_result(self) -> Value:
        out, e = CValue(), _err()
        if lib().orc_fn_accumulate_result(self._h, C.byref(out), e):
            _raise(e)
        v = _from_cvalue(out)
        lib().orc_value_free(C.byref(out))
        return v

    def merge_state(self, states: Value):
        keep: list = []
        cv = _to_cvalue(states, keep)
        e = _err()
        if lib().orc_fn_merge_state(self._h, C.byref(cv), e):
            _raise(e)

    def merge_result(self) -> Value:
        out, e = CValue(), _err()
        if lib().orc_fn_merge_result(self._h, C.byref(out), e):
            _raise(e)
        v = _from_cvalue(out)
        lib().orc_value_free(C.byref(out))
        return v


# ---------------------------------------------------------------------------------------------
# pipeline
# ---------------------------------------------------------------------------------------------
@dataclass
class QueryResult:
    names: List[str]
    columns: List[Array]
    n_rows: int
    n_blocks_out: int
    rows_scanned: int
    partial_states_json: List[str] = field(default_factory=list)
    seconds: float = 0.0

    def rows(self):
        cols = [c.to_list() for c in self.columns]
        return [tuple(c[i] for c in cols) for i in range(self.n_rows)]


def fused_headline(total: int, threads: int):
    """Best-case CPU pass of the headline query (not the reference's structure): (seconds, [sum, count, max, min])."""
    L = lib()
    L.orc_fused_headline.argtypes = [C.c_uint64, C.c_int32, C.POINTER(C.c_uint64)]
    L.orc_fused_headline.restype = C.c_double
    out = (C.c_uint64 * 4)()
    secs = L.orc_fused_headline(total, threads, out)
    return secs, [int(x) for x in out]


def run_query(exprs: Sequence[str], *, total: int = 10000, table: Optional[dict] = None, predicate: Optional[str] = None,
              is_aggregate: bool = False, limit: Optional[int] = None, worker_threads: int = 8, use_threads: bool = False,
              block_size: int = 10000, tail_quirk: bool = True) -> QueryResult:
    """Run the reference-shaped pipeline.  `exprs`/`predicate` are s-expressions (see fq_oracle.h)."""
    keep: list = []
    q = CQuery()
    q.total = total
    if table is not None:
        tb = _block(table, keep)
        keep.append(tb)
        q.table = C.pointer(tb)
    q.block_size = block_size
    q.tail_quirk = int(tail_quirk)
    q.worker_threads = worker_threads
    q.use_threads = int(use_threads)
    q.predicate = predicate.encode() if predicate else None
    q.n_exprs = len(exprs)
    q.is_aggregate = int(is_aggregate)
    arr = (C.c_char_p * len(exprs))(*[s.encode() for s in exprs])
    q.exprs = C.cast(arr, C.POINTER(C.c_char_p))
    q.limit = -1 if limit is None else limit
    res, e = CResult(), _err()
    if lib().orc_query_run(C.byref(q), C.byref(res), e):
        _raise(e)
    names = [res.block.names[i].decode() for i in range(res.block.n_cols)]
    cols = [_from_carray(res.block.cols[i]) for i in range(res.block.n_cols)]
    js = C.string_at(res.partial_states_json).decode().split("\n") if res.partial_states_json else []
    out = QueryResult(names, cols, res.n_rows, res.n_blocks_out, res.rows_scanned, js, res.seconds)
    lib().orc_result_free(C.byref(res))
    return out
