"""The oracle's arrow semantics against an independent Arrow implementation.

The reference's arithmetic lives in the arrow crate 2.0.0, which is not under /root/reference.  The
oracle restates it; the reference's own test vectors pin the non-null cases (tests/test_oracle_golden.py).  This file
adds a second anchor for what those vectors do not cover — NULL propagation, wrapping at the type's width, aggregates
over NULLs — by replaying random arrays through pyarrow.compute (Arrow C++), whose kernels implement the same
specification.  Where Arrow C++ and arrow-rs 2.0.0 are known to differ (out-of-range `cast`: C++ raises, Rust yields
NULL; same-type only here) the case is left to the oracle's own tests."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from oracle import binding as o

INT_TYPES = [(o.I8, np.int8), (o.I16, np.int16), (o.I32, np.int32), (o.I64, np.int64), (o.U8, np.uint8), (o.U16, np.uint16),
             (o.U32, np.uint32), (o.U64, np.uint64)]
FLOAT_TYPES = [(o.F32, np.float32), (o.F64, np.float64)]
N = 4099


def make(rng, npdt, nonzero=False, null_frac=0.3):
    if np.issubdtype(npdt, np.integer):
        info = np.iinfo(npdt)
        v = rng.integers(info.min, info.max, N, dtype=npdt, endpoint=True)
    else:
        v = (rng.normal(0, 1e3, N)).astype(npdt)
    valid = (rng.random(N) > null_frac).astype(np.uint8)
    if nonzero:
        v = np.where((v == 0) & (valid == 1), 1, v).astype(npdt)
        v[valid == 0] = 0                      # zero divisors hide in NULL slots only
    return v, valid


def as_arrow(v, valid):
    return pa.array(v, mask=(valid == 0))


def assert_same(got: o.Array, want: pa.Array, exact_float=True):
    gv = np.ones(len(got), np.uint8) if got.valid is None else got.valid
    wv = np.asarray(want.is_valid()).astype(np.uint8)
    assert np.array_equal(gv, wv)
    w = want.fill_null(False if pa.types.is_boolean(want.type) else 0).to_numpy(zero_copy_only=False)
    m = wv.astype(bool)
    g = got.values[m]
    assert np.array_equal(g.astype(w.dtype), w[m], equal_nan=True)


@pytest.mark.parametrize("tag,npdt", INT_TYPES + FLOAT_TYPES)
@pytest.mark.parametrize("op,fn", [("+", pc.add), ("-", pc.subtract), ("*", pc.multiply)])
def test_arithmetic_wraps_and_propagates_nulls(tag, npdt, op, fn):
    rng = np.random.default_rng(hash((tag, op)) & 0xffff)
    (a, av), (b, bv) = make(rng, npdt), make(rng, npdt)
    got = o.array_arithmetic(op, o.array(tag, a, av), o.array(tag, b, bv))
    assert got.dtype == tag
    assert_same(got, fn(as_arrow(a, av), as_arrow(b, bv)))      # unchecked kernels: integers wrap at the type's width


@pytest.mark.parametrize("tag,npdt", INT_TYPES)
def test_integer_divide_truncates_and_ignores_null_slots(tag, npdt):
    rng = np.random.default_rng(tag)
    (a, av), (b, bv) = make(rng, npdt), make(rng, npdt, nonzero=True)
    if np.issubdtype(npdt, np.signedinteger):
        a = np.where(a == np.iinfo(npdt).min, 0, a).astype(npdt)   # MIN / -1 overflows: unpinned, kept out
    got = o.array_arithmetic("/", o.array(tag, a, av), o.array(tag, b, bv))
    assert_same(got, pc.divide(as_arrow(a, av), as_arrow(b, bv)))
    b2 = b.copy()
    b2[np.flatnonzero(bv & av)[0]] = 0                         # one zero divisor in a slot both sides hold valid
    with pytest.raises(o.OracleError) as e:
        o.array_arithmetic("/", o.array(tag, a, av), o.array(tag, b2, bv))
    assert str(e.value) == "Internal Error: Divide by zero error"
    with pytest.raises(pa.ArrowInvalid):
        pc.divide(as_arrow(a, av), as_arrow(b2, bv))


@pytest.mark.parametrize("tag,npdt", INT_TYPES + FLOAT_TYPES)
@pytest.mark.parametrize("op,fn", [("=", pc.equal), ("<", pc.less), ("<=", pc.less_equal), (">", pc.greater), (">=", pc.greater_equal)])
def test_comparisons_propagate_nulls(tag, npdt, op, fn):
    rng = np.random.default_rng(hash((tag, op)) & 0xffff)
    (a, av), (b, bv) = make(rng, npdt), make(rng, npdt)
    b[::7] = a[::7]                                            # some equal pairs
    got = o.array_comparison(op, o.array(tag, a, av), o.array(tag, b, bv))
    assert got.dtype == o.BOOL
    assert_same(got, fn(as_arrow(a, av), as_arrow(b, bv)))


@pytest.mark.parametrize("op,fn", [("and", pc.and_), ("or", pc.or_)])
def test_logic_is_not_kleene(op, fn):
    rng = np.random.default_rng(5)
    a, b = rng.integers(0, 2, N).astype(np.uint8), rng.integers(0, 2, N).astype(np.uint8)
    av, bv = (rng.random(N) > 0.3).astype(np.uint8), (rng.random(N) > 0.3).astype(np.uint8)
    got = o.array_logic(op, o.array(o.BOOL, a, av), o.array(o.BOOL, b, bv))
    assert_same(got, fn(pa.array(a.astype(bool), mask=av == 0), pa.array(b.astype(bool), mask=bv == 0)))   # and_/or_: NULL if either is


@pytest.mark.parametrize("tag,npdt", INT_TYPES + FLOAT_TYPES)
def test_aggregates_skip_nulls_and_empty_is_none(tag, npdt):
    rng = np.random.default_rng(100 + tag)
    a, av = make(rng, npdt)
    if np.issubdtype(npdt, np.floating):
        a = np.round(a).astype(npdt)                           # exact float sums whatever the order
    arr, ref = o.array(tag, a, av), as_arrow(a, av)
    assert o.array_aggregate("min", arr).value == pc.min(ref).as_py()
    assert o.array_aggregate("max", arr).value == pc.max(ref).as_py()
    if np.issubdtype(npdt, np.integer):
        want = int(a[av == 1].astype(np.uint64 if np.issubdtype(npdt, np.unsignedinteger) else np.int64).sum(dtype=npdt))   # wraps at the type's width
        assert o.array_aggregate("sum", arr).value == want
    else:
        assert o.array_aggregate("sum", arr).value == pytest.approx(pc.sum(ref).as_py(), rel=1e-6)
    assert o.array_aggregate("count", arr).value == N          # the reference's Count is the array length (data_array_aggregate.rs:29)
    none = o.array(tag, a, np.zeros(N, np.uint8))
    for op, fn in (("sum", pc.sum), ("min", pc.min), ("max", pc.max)):
        assert o.array_aggregate(op, none).value is None and fn(pa.array(a, mask=np.ones(N, bool))).as_py() is None
    empty = o.array(tag, a[:0])
    for op in ("sum", "min", "max"):
        assert o.array_aggregate(op, empty).value is None
