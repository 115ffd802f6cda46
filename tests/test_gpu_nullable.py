"""Validity through the fused kernels (the data formats either side of the path).

The reference's DataArrayRef is an arrow ArrayRef, so every kernel it calls carries a null bitmap
(data_array_arithmetic.rs:16-50 -> arrow compute::add/..., data_array_comparison.rs:16-53,
data_array_logic.rs:15-41, data_array_aggregate.rs:18-80) and `cast` turns out-of-range values into
NULL slots (numerical_arithmetic_coercion, data_array_arithmetic.rs:24-27).  system.numbers never
produces a NULL, so the reference's own tests hold no vector for this; the oracle restates arrow
2.0.0's published semantics and these tests hold the device to the oracle.

Values are compared only in valid slots (arrow leaves the payload of a NULL slot unspecified)."""
import random

import numpy as np
import pytest

from fuse_query_b200 import cabi
from oracle import binding as o

pytestmark = pytest.mark.gpu

N = 70_003
NAMES = ["a", "b", "c", "e"]
DT = {"a": cabi.U64, "b": cabi.I64, "c": cabi.I32, "e": cabi.F64}


@pytest.fixture(scope="module")
def env():
    ctx = cabi.Context(0)
    rng = np.random.default_rng(20260)
    vals = {
        "a": rng.integers(1, 1 << 40, N, dtype=np.uint64),
        "b": rng.integers(-(1 << 40), 1 << 40, N, dtype=np.int64),
        "c": rng.integers(1, 30000, N, dtype=np.int32),
        "e": np.round(rng.normal(0, 1e6, N), 3),
    }
    valid = {
        "a": (rng.random(N) > 0.2).astype(np.uint8),
        "b": (rng.random(N) > 0.5).astype(np.uint8),
        "c": None,                                   # NOT NULL next to nullable columns
        "e": (rng.random(N) > 0.05).astype(np.uint8),
    }
    cols = [ctx.from_numpy(vals[k], valid[k]) for k in NAMES]
    table = {k: o.array(DT[k], vals[k], valid[k]) for k in NAMES}
    yield ctx, vals, valid, cols, table
    ctx.close()


def project(ctx, cols, exprs, pred=None, n=N, nullable=None, names=NAMES, dtypes=None):
    nullable = [c.validity is not None for c in cols] if nullable is None else nullable
    pipe = ctx.pipe(exprs, columns=names, dtypes=dtypes or [DT[k] for k in names], predicate=pred, nullable=nullable)
    outs = [ctx.column(pipe.expr_dtype(i), n) for i in range(len(exprs))]
    ov = [ctx.column(cabi.BOOL, n) if pipe.expr_nullable(i) else None for i in range(len(exprs))]
    pipe.launch_project(cabi.make_source(cols, n), outs, n, out_valid=ov)
    sel, written = pipe.fetch_project()
    assert sel == written
    return pipe, [c.to_numpy(written) for c in outs], [None if v is None else v.to_numpy(written) for v in ov], written


def check(pipe, got, got_valid, want, exprs):
    for i, c in enumerate(want.columns):
        assert pipe.expr_dtype(i) == c.dtype, exprs[i]
        wv = np.ones(len(c.values), np.uint8) if c.valid is None else c.valid
        gv = np.ones(len(got[i]), np.uint8) if got_valid[i] is None else got_valid[i].astype(np.uint8)
        assert np.array_equal(gv, wv), f"validity of {exprs[i]}"
        m = wv.astype(bool)
        g, w = got[i][m], c.values[m]
        if c.dtype in (o.F32, o.F64):
            assert np.array_equal(g, w, equal_nan=True), exprs[i]
        else:
            assert np.array_equal(g.astype(w.dtype), w), exprs[i]


PROJECTIONS = [
    ["(+ (col a) (u64 1))", "(* (col a) (col a))", "(col a)"],
    ["(+ (col a) (col b))", "(- (col b) (col c))", "(/ (col b) (col c))"],
    ["(< (col a) (u64 500000000000))", "(= (col b) (col b))", "(and (> (col b) (i64 0)) (< (col a) (u64 900000000000)))"],
    ["(* (col e) (f64 2.5))", "(+ (col e) (col c))", "(or (< (col e) (f64 0)) (> (col c) (i32 100)))"],
    ["(col c)", "(+ (col c) (i32 7))", "(/ (col a) (col c))"],          # a NOT NULL column stays NOT NULL
]


@pytest.mark.parametrize("exprs", PROJECTIONS)
def test_projection_over_nullable_columns(env, exprs):
    ctx, vals, valid, cols, table = env
    want = o.run_query(exprs, table=table, worker_threads=1, tail_quirk=False)
    pipe, got, gv, written = project(ctx, cols, exprs)
    assert written == want.n_rows == N
    check(pipe, got, gv, want, exprs)


def test_not_null_expression_reports_not_nullable(env):
    ctx, vals, valid, cols, table = env
    pipe = ctx.pipe(["(+ (col c) (i32 7))", "(+ (col a) (col c))"], columns=NAMES, dtypes=[DT[k] for k in NAMES],
                    nullable=[True, True, False, True])
    assert pipe.expr_nullable(0) is False and pipe.expr_nullable(1) is True


@pytest.mark.parametrize("pred", [
    "(> (col b) (i64 0))",                                   # NULL predicate slots keep nothing
    "(and (< (col a) (u64 700000000000)) (>= (col e) (f64 -500000)))",
    "(or (= (col c) (i32 5)) (> (col b) (col c)))",
])
def test_filter_with_nullable_predicate(env, pred):
    ctx, vals, valid, cols, table = env
    exprs = ["(col a)", "(+ (col b) (i64 1))", "(col c)"]
    want = o.run_query(exprs, table=table, predicate=pred, worker_threads=1, tail_quirk=False)
    pipe, got, gv, written = project(ctx, cols, exprs, pred)
    assert 0 < written == want.n_rows < N
    check(pipe, got, gv, want, exprs)


@pytest.mark.parametrize("pred", [None, "(> (col c) (i32 15000))", "(> (col b) (i64 0))"])
def test_aggregates_skip_null_slots(env, pred):
    ctx, vals, valid, cols, table = env
    exprs = ["(sum (col a))", "(min (col b))", "(max (col e))", "(count (col b))", "(sum (+ (col a) (col b)))", "(count (col c))"]
    want = [o.run_query([e], table=table, predicate=pred, is_aggregate=True, worker_threads=1, tail_quirk=False,
                        block_size=1 << 30).columns[0] for e in exprs]
    pipe = ctx.pipe(exprs, columns=NAMES, dtypes=[DT[k] for k in NAMES], predicate=pred, aggregate=True,
                    nullable=[c.validity is not None for c in cols])
    pipe.launch_aggregate(cabi.make_source(cols, N))
    states, rows = pipe.fetch_aggregate()
    for (dtype, val), w, e in zip(states, want, exprs):
        assert dtype == w.dtype, e
        assert val == w.to_list()[0], e


def test_aggregate_over_all_null_column_is_none(env):
    ctx, vals, valid, cols, table = env
    n = 4096
    x = np.arange(n, dtype=np.uint64)
    col = ctx.from_numpy(x, np.zeros(n, np.uint8))
    pipe = ctx.pipe(["(sum (col x))", "(min (col x))", "(max (col x))", "(count (col x))"], columns=["x"], dtypes=[cabi.U64],
                    aggregate=True, nullable=[True])
    pipe.launch_aggregate(cabi.make_source([col], n))
    states, rows = pipe.fetch_aggregate()
    assert [v for _, v in states[:3]] == [None, None, None]
    # the oracle agrees, leaf by leaf (arrow sum/min/max over an all-NULL array is None; Count is the array length, NULL slots included: data_array_aggregate.rs:29)
    t = {"x": o.array(o.U64, x, np.zeros(n, np.uint8))}
    for op, (_, v) in zip(["sum", "min", "max", "count"], states):
        assert o.array_aggregate(op, t["x"]).value == v, op


def test_division_by_zero_in_a_null_slot_is_not_an_error(env):
    """arrow's divide looks at the divisor only where both operands are valid (arithmetic.rs math_divide)."""
    ctx, *_ = env
    n = 10_000
    num = np.arange(n, dtype=np.uint64) + 10
    den = (np.arange(n, dtype=np.uint64) % 7)           # zero every 7th row
    den_valid = (den != 0).astype(np.uint8)             # ... and exactly those rows are NULL
    cols = [ctx.from_numpy(num), ctx.from_numpy(den, den_valid)]
    table = {"x": o.array(o.U64, num), "y": o.array(o.U64, den, den_valid)}
    exprs = ["(/ (col x) (col y))"]
    want = o.run_query(exprs, table=table, worker_threads=1, tail_quirk=False)
    pipe, got, gv, written = project(ctx, cols, exprs, n=n, names=["x", "y"], dtypes=[cabi.U64, cabi.U64])
    check(pipe, got, gv, want, exprs)
    # the same divisor declared NOT NULL does divide by zero
    cols2 = [ctx.from_numpy(num), ctx.from_numpy(den)]
    with pytest.raises(cabi.FuseGpuError) as ei:
        project(ctx, cols2, exprs, n=n, names=["x", "y"], dtypes=[cabi.U64, cabi.U64])
    with pytest.raises(o.OracleError) as eo:
        o.run_query(exprs, table={"x": o.array(o.U64, num), "y": o.array(o.U64, den)}, worker_threads=1, tail_quirk=False)
    assert str(ei.value) == str(eo.value)


def test_out_of_range_cast_yields_null(env):
    """UInt64 (+) Int64 coerces to Int64 (data_type.rs:27-98); a UInt64 above i64::MAX casts to NULL."""
    ctx, *_ = env
    n = 33_333
    rng = np.random.default_rng(5)
    a = rng.integers(0, 1 << 64, n, dtype=np.uint64)     # about half above i64::MAX
    b = rng.integers(-1000, 1000, n, dtype=np.int64)
    s = rng.integers(-100, 100, n).astype(np.int8)
    cols = [ctx.from_numpy(a), ctx.from_numpy(b), ctx.from_numpy(s)]
    table = {"a": o.from_numpy(a), "b": o.from_numpy(b), "s": o.from_numpy(s)}
    exprs = ["(+ (col a) (col b))", "(< (col a) (col b))", "(* (col s) (col a))", "(+ (col a) (i64 -3))"]
    want = o.run_query(exprs, table=table, worker_threads=1, tail_quirk=False)
    assert any(c.valid is not None and not c.valid.all() for c in want.columns)
    pipe, got, gv, written = project(ctx, cols, exprs, n=n, names=["a", "b", "s"], dtypes=[cabi.U64, cabi.I64, cabi.I8])
    check(pipe, got, gv, want, exprs)
    # ... and the aggregates over such an expression skip the NULL slots
    aggs = ["(sum (+ (col a) (col b)))", "(count (+ (col a) (col b)))", "(min (+ (col a) (col b)))"]
    wanta = [o.run_query([e], table=table, is_aggregate=True, worker_threads=1, tail_quirk=False, block_size=1 << 30).columns[0]
             for e in aggs]
    pipe = ctx.pipe(aggs, columns=["a", "b", "s"], dtypes=[cabi.U64, cabi.I64, cabi.I8], aggregate=True)
    pipe.launch_aggregate(cabi.make_source(cols, n))
    states, _ = pipe.fetch_aggregate()
    for (dtype, val), w, e in zip(states, wanta, aggs):
        assert dtype == w.dtype and val == w.to_list()[0], e


def test_sliced_nullable_column_keeps_its_validity(env):
    ctx, vals, valid, cols, table = env
    off, n = 12_345, 20_001     # off the 16-byte grid: the kernels fall back to row-by-row loads
    sl = [c.slice(off, n) for c in cols]
    exprs = ["(+ (col a) (col b))", "(col e)"]
    t2 = {k: o.array(DT[k], vals[k][off:off + n], None if valid[k] is None else valid[k][off:off + n]) for k in NAMES}
    want = o.run_query(exprs, table=t2, worker_threads=1, tail_quirk=False)
    pipe, got, gv, written = project(ctx, sl, exprs, n=n, nullable=[valid[k] is not None for k in NAMES])
    check(pipe, got, gv, want, exprs)


def test_validity_mismatch_is_rejected(env):
    ctx, vals, valid, cols, table = env
    plain = ctx.from_numpy(vals["a"])
    pipe = ctx.pipe(["(col a)"], columns=["a"], dtypes=[cabi.U64], nullable=[True])
    out = ctx.column(cabi.U64, N)
    with pytest.raises(cabi.FuseGpuError):
        pipe.launch_project(cabi.make_source([plain], N), [out], N, out_valid=[ctx.column(cabi.BOOL, N)])


@pytest.mark.parametrize("seed", range(12))
def test_random_trees_over_nullable_columns(env, seed):
    ctx, vals, valid, cols, table = env
    rng = random.Random(1000 + seed)

    def num(d):
        if d == 0 or rng.random() < 0.3:
            if rng.random() < 0.75:
                return f"(col {rng.choice(NAMES)})"
            return rng.choice([f"(u64 {rng.randint(1, 10**6)})", f"(i64 {rng.randint(-10**6, 10**6)})", f"(f64 {round(rng.uniform(-9, 9), 2)})"])
        op = rng.choice(["+", "-", "*", "/"])
        l = num(d - 1)
        r = rng.choice(["(col c)", f"(u64 {rng.randint(1, 99)})"]) if op == "/" else num(d - 1)
        if not l.startswith("(col") and not r.startswith("(col") and "(col" not in l + r:
            l = f"(col {rng.choice(NAMES)})"
        return f"({op} {l} {r})"

    def pred(d):
        if d and rng.random() < 0.5:
            return f"({rng.choice(['and', 'or'])} {pred(d - 1)} {pred(d - 1)})"
        return f"({rng.choice(['=', '<', '<=', '>', '>='])} {num(1)} {num(1)})"

    exprs = [num(3), num(2), pred(1)]
    p = pred(1) if seed % 2 else None
    try:
        want = o.run_query(exprs, table=table, predicate=p, worker_threads=1, tail_quirk=False)
    except o.OracleError as e:
        with pytest.raises(cabi.FuseGpuError) as ei:
            project(ctx, cols, exprs, p)
        assert str(ei.value) == str(e)
        return
    pipe, got, gv, written = project(ctx, cols, exprs, p)
    assert written == want.n_rows
    check(pipe, got, gv, want, exprs)


@pytest.mark.parametrize("n,bit_offset", [(1, 0), (7, 0), (8, 0), (9, 3), (64, 5), (1000, 7), (100_003, 0), (100_003, 13), (1_000_001, 6)])
def test_arrow_bitmaps_expand_and_pack_on_the_device(env, n, bit_offset):
    """include/fuse_gpu.h fq_column_upload_bits / _download_bits: arrow's LSB-first validity / Boolean bitmaps
    (arrow 2.0.0 bitmap.rs) <-> one byte per row, against numpy's packbits / unpackbits (bitorder='little')."""
    ctx, *_ = env
    rng = np.random.default_rng(n + bit_offset)
    bits = rng.integers(0, 256, (bit_offset + n + 7) // 8 + 2, dtype=np.uint8)
    want = np.unpackbits(bits, bitorder="little")[bit_offset:bit_offset + n]
    col = ctx.from_bitmap(bits, n, bit_offset)
    assert np.array_equal(col.to_numpy(n), want)
    packed = col.to_bitmap(n)
    assert np.array_equal(packed, np.packbits(want, bitorder="little"))
    # a slice that starts off the byte grid packs from its own row 0
    if n > 11:
        sl = col.slice(3, n - 11)
        assert np.array_equal(sl.to_bitmap(), np.packbits(want[3:n - 8], bitorder="little"))


def test_pyarrow_validity_bitmap_goes_to_the_device_unexpanded(env):
    """tables.register_table hands arrow's own null bitmap (with the array's offset) to the device."""
    import pyarrow as pa
    from fuse_query_b200 import _fuse_host as h
    from fuse_query_b200.tables import register_table
    gpu = h.GpuContext.create(0)
    fctx = h.FuseQueryContext.create_ctx(1, gpu)
    n = 10_007
    rng = np.random.default_rng(1)
    vals = rng.integers(0, 1000, n)
    mask = rng.random(n) < 0.3
    arr = pa.array(vals, type=pa.int64(), mask=mask).slice(5, n - 9)      # a non-zero arrow offset
    t = register_table(fctx, gpu, "default", "bm", {"x": arr, "y": np.arange(n - 9, dtype=np.uint64)})
    assert t.schema().fields[0].nullable
    got = h.execute_sql(fctx, "select x, y from bm")[0]
    want = [None if m else int(v) for v, m in zip(vals[5:n - 4], mask[5:n - 4])]
    assert got.column(0).to_list() == want
    assert got.column(0).validity().to_arrow_bitmap() == np.packbits(~mask[5:n - 4], bitorder="little").tobytes()
    (cnt,) = h.execute_sql(fctx, "select sum(x) from bm")[0].column(0).to_list()
    assert cnt == int(vals[5:n - 4][~mask[5:n - 4]].sum())


@pytest.mark.parametrize("variant", ["tma", "ldg"])
def test_arrow_validity_bitmaps_are_read_in_place(env, variant):
    """fq_column_set_validity_bitmap: Arrow's LSB-first validity buffer stays bit-packed on the device (pipes compiled with
    col_nullable = 2) — 1 bit of validity traffic per row instead of 1 byte.  Same results as the byte-per-row form and as the
    oracle, through every kernel family (aggregate, filter + projection, projection, GROUP BY), with a bit offset (arrow's
    array offset), on slices, and when a thread's bits straddle a byte (offsets that are not multiples of the vector width)."""
    ctx, vals, valid, _cols, table = env
    os_env = __import__("os").environ
    for name, var in (("FQ_AGG_VARIANT", "tma" if variant == "tma" else "u4"), ("FQ_SEL_VARIANT", variant), ("FQ_MAP_VARIANT", variant)):
        os_env[name] = var
    try:
        for bit_offset in (0, 3, 64):
            cols = []
            for k in NAMES:
                if valid[k] is None:
                    cols.append(ctx.from_numpy(vals[k]))
                else:
                    bits = np.packbits(np.concatenate([np.zeros(bit_offset, np.uint8), valid[k]]), bitorder="little")
                    cols.append(ctx.from_numpy(vals[k], valid_bitmap=bits, bit_offset=bit_offset))
            nullable = [0 if valid[k] is None else 2 for k in NAMES]
            kw = dict(columns=NAMES, dtypes=[DT[k] for k in NAMES], nullable=nullable)
            # aggregates
            aggs = ["(sum (col a))", "(min (col b))", "(max (+ (col b) (col c)))", "(count (col a))", "(sum (col e))"]
            pipe = ctx.pipe(aggs, aggregate=True, predicate="(> (col c) (i32 100))", **kw)
            pipe.launch_aggregate(cabi.make_source(cols, N))
            states, rows = pipe.fetch_aggregate()
            want = o.run_query(aggs, table=table, predicate="(> (col c) (i32 100))", is_aggregate=True, worker_threads=1, tail_quirk=False,
                               block_size=1 << 30)
            got = [s[1] for s in states]
            exp = list(want.rows()[0])
            assert got[:4] == exp[:4] and abs(got[4] - exp[4]) <= 1e-9 * abs(exp[4])
            # filter + projection, on a slice that starts off the vector grid (rows 5 .. N-2)
            exprs = ["(+ (col a) (u64 1))", "(* (col b) (col c))"]
            pred = "(< (col b) (i64 0))"
            for off, m in ((0, N), (5, N - 7), (64, 10_000)):
                sl = [c.slice(off, m) for c in cols]
                p2 = ctx.pipe(exprs, predicate=pred, **kw)
                outs = [ctx.column(p2.expr_dtype(i), m) for i in range(2)]
                ov = [ctx.column(cabi.BOOL, m) if p2.expr_nullable(i) else None for i in range(2)]
                p2.launch_project(cabi.make_source(sl, m), outs, m, out_valid=ov)
                sel, written = p2.fetch_project()
                sub = {k: o.array(DT[k], vals[k][off:off + m], None if valid[k] is None else valid[k][off:off + m]) for k in NAMES}
                w = o.run_query(exprs, table=sub, predicate=pred, worker_threads=1, tail_quirk=False, block_size=1 << 30)
                assert written == w.n_rows
                for i in range(2):
                    wv = w.columns[i]
                    ok = np.ones(written, bool) if wv.valid is None else wv.valid.astype(bool)
                    gv = np.ones(written, bool) if ov[i] is None else ov[i].to_numpy(written).astype(bool)
                    assert np.array_equal(gv, ok) and np.array_equal(outs[i].to_numpy(written)[ok], wv.values[ok])
                p2.destroy()
            pipe.destroy()
        # a pipe compiled for bitmaps refuses a byte-validity column and vice versa
        pipe = ctx.pipe(["(sum (col a))"], aggregate=True, columns=["a"], dtypes=[cabi.U64], nullable=[2])
        with pytest.raises(cabi.FuseGpuError) as e:
            pipe.launch_aggregate(cabi.make_source([ctx.from_numpy(vals["a"], valid["a"])], N))
        assert "carries no validity bitmap" in str(e.value)
    finally:
        for name in ("FQ_AGG_VARIANT", "FQ_SEL_VARIANT", "FQ_MAP_VARIANT"):
            os_env.pop(name, None)
