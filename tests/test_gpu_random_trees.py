"""Randomly generated expression trees (seeded) through NVRTC, against the oracle: arithmetic / comparison /
logic over mixed-type columns and literals, as projections, predicates and aggregate arguments.
Integer and boolean results bit-exact; Float64 element-wise exact (same IEEE ops), float sums 1e-9 relative."""
import random

import numpy as np
import pytest

from fuse_query_b200 import cabi
from oracle import binding as o

pytestmark = pytest.mark.gpu

COLS = {"a": (cabi.U64, np.uint64), "b": (cabi.I64, np.int64), "c": (cabi.I32, np.int32), "d": (cabi.U16, np.uint16),
        "e": (cabi.F64, np.float64), "f": (cabi.U8, np.uint8)}
NAMES = list(COLS)
N = 50_021


@pytest.fixture(scope="module")
def env():
    ctx = cabi.Context(0)
    rng = np.random.default_rng(12345)
    tbl = {
        "a": rng.integers(0, 1 << 40, N, dtype=np.uint64),
        "b": rng.integers(-(1 << 40), 1 << 40, N, dtype=np.int64),
        "c": rng.integers(-30000, 30000, N, dtype=np.int32),
        "d": rng.integers(1, 60000, N, dtype=np.uint16),
        "e": np.round(rng.normal(0, 1e6, N), 3),
        "f": rng.integers(1, 200, N, dtype=np.uint8),
    }
    cols = [ctx.from_numpy(tbl[k]) for k in NAMES]
    yield ctx, tbl, cols
    ctx.close()


def gen_num(rng, depth):
    """numeric expression s-expr; divisors are kept non-zero (columns d, f are >= 1; literals >= 1)"""
    if depth == 0 or rng.random() < 0.25:
        if rng.random() < 0.7:
            return f"(col {rng.choice(NAMES)})"
        kind = rng.choice(["u64", "i64", "u8", "i16", "f64"])
        lit = {"u64": rng.randint(0, 10**6), "i64": rng.randint(-10**6, 10**6), "u8": rng.randint(1, 200),
               "i16": rng.randint(-300, 300), "f64": round(rng.uniform(-100, 100), 2)}[kind]
        return f"({kind} {lit})"
    op = rng.choice(["+", "-", "*", "/", "+", "-"])
    l = gen_num(rng, depth - 1)
    if op == "/":
        r = rng.choice(["(col d)", "(col f)", f"(u64 {rng.randint(1, 1000)})", f"(i64 {rng.randint(1, 1000)})"])
    else:
        r = gen_num(rng, depth - 1)
    if l.startswith(("(u64", "(i64", "(u8", "(i16", "(f64")) and r.startswith(("(u64", "(i64", "(u8", "(i16", "(f64")):
        l = f"(col {rng.choice(NAMES)})"   # constant (op) constant is a 1-row array in the reference: not a row expression
    return f"({op} {l} {r})"


def gen_pred(rng, depth):
    if depth > 0 and rng.random() < 0.4:
        return f"({rng.choice(['and', 'or'])} {gen_pred(rng, depth - 1)} {gen_pred(rng, depth - 1)})"
    return f"({rng.choice(['=', '<', '<=', '>', '>='])} {gen_num(rng, 1)} {gen_num(rng, 1)})"


def oracle_or_error(fn):
    try:
        return fn(), None
    except o.OracleError as e:
        return None, str(e)


@pytest.mark.parametrize("seed", range(24))
def test_random_projection_and_filter(env, seed):
    ctx, tbl, cols = env
    rng = random.Random(seed)
    exprs = [gen_num(rng, 3), gen_num(rng, 2), gen_pred(rng, 1)]
    pred = gen_pred(rng, 2) if seed % 2 else None
    table = {k: o.from_numpy(v) for k, v in tbl.items()}
    want, err = oracle_or_error(lambda: o.run_query(exprs, table=table, predicate=pred, worker_threads=1, tail_quirk=False))
    try:
        pipe = ctx.pipe(exprs, columns=NAMES, dtypes=[COLS[k][0] for k in NAMES], predicate=pred)
        outs = [ctx.column(pipe.expr_dtype(i), N) for i in range(len(exprs))]
        # an expression holding a fallible cast (say UInt64 -> Int64) can yield NULL slots: it gets a validity output
        ov = [ctx.column(cabi.BOOL, N) if pipe.expr_nullable(i) else None for i in range(len(exprs))]
        pipe.launch_project(cabi.make_source(cols, N), outs, N, out_valid=ov)
        sel, written = pipe.fetch_project()
    except cabi.FuseGpuError as e:
        assert err is not None and str(e) == err, (exprs, pred, str(e), err)
        return
    assert err is None, (exprs, pred, err)
    assert written == sel == want.n_rows
    for i, c in enumerate(want.columns):
        got = outs[i].to_numpy(written)
        assert pipe.expr_dtype(i) == c.dtype, (exprs[i], pipe.expr_dtype(i), c.dtype)
        exp = c.values
        if c.valid is not None or ov[i] is not None:
            wv = np.ones(written, np.uint8) if c.valid is None else c.valid
            gv = np.ones(written, np.uint8) if ov[i] is None else ov[i].to_numpy(written)
            assert np.array_equal(gv, wv), f"validity of {exprs[i]}"
            got, exp = got[wv.astype(bool)], exp[wv.astype(bool)]
        if c.dtype in (o.F32, o.F64):
            assert np.array_equal(got, exp, equal_nan=True), exprs[i]
        else:
            assert np.array_equal(got.astype(exp.dtype), exp), exprs[i]


@pytest.mark.parametrize("seed", range(100, 116))
def test_random_aggregates(env, seed):
    ctx, tbl, cols = env
    rng = random.Random(seed)
    ops = ["sum", "min", "max", "count"]
    exprs = [f"({rng.choice(ops)} {gen_num(rng, 2)})" for _ in range(3)]
    pred = gen_pred(rng, 1) if seed % 3 == 0 else None
    table = {k: o.from_numpy(v) for k, v in tbl.items()}
    want = []
    err = None
    for e in exprs:
        # one big block: the reference's per-10 000-row-block Sum poisoning is a block artefact, not arithmetic
        r, er = oracle_or_error(lambda: o.run_query([e], table=table, predicate=pred, is_aggregate=True, worker_threads=1,
                                                    tail_quirk=False, block_size=1 << 30))
        if er == "Internal Error: DataValue to array cannot be NONE NULL":
            # Sum/Min/Max over zero selected rows is Type(None); the reference then fails in AggregateFinalTransform's
            # merge_result().to_array(1) (transform_aggregate_final.rs:68-72).  The device must report None for the leaf.
            want.append((None, None))
            continue
        err = err or er
        want.append(None if r is None else (r.columns[0].to_list()[0], r.columns[0].dtype))
    try:
        pipe = ctx.pipe(exprs, columns=NAMES, dtypes=[COLS[k][0] for k in NAMES], predicate=pred, aggregate=True)
        pipe.launch_aggregate(cabi.make_source(cols, N))
        states, rows = pipe.fetch_aggregate()
    except cabi.FuseGpuError as e:
        assert err is not None and str(e) == err, (exprs, pred, str(e), err)
        return
    assert err is None, (exprs, err)
    for (dtype, val), (w, wt), e in zip(states, want, exprs):
        if wt is None:   # no selected row, or every selected row NULL (an out-of-range cast inside the argument)
            assert val is None, e
            continue
        assert dtype == wt, e
        if dtype in (cabi.F32, cabi.F64) and e.startswith("(sum"):
            assert val == pytest.approx(w, rel=1e-9, abs=1e-6), e
        else:
            assert val == w, e


# ---------------------------------------------------------------------------------------------
# the narrow and single-precision lanes of the coercion lattice (data_type.rs:27-98): Int8, Int16, UInt32, Float32
# ---------------------------------------------------------------------------------------------
COLS2 = {"p": (cabi.I8, np.int8), "q": (cabi.I16, np.int16), "r": (cabi.U32, np.uint32), "s": (cabi.F32, np.float32),
         "t": (cabi.U8, np.uint8), "u": (cabi.I64, np.int64)}
NAMES2 = list(COLS2)


@pytest.fixture(scope="module")
def env2():
    ctx = cabi.Context(0)
    rng = np.random.default_rng(777)
    tbl = {
        "p": rng.integers(-128, 128, N).astype(np.int8),
        "q": rng.integers(-32768, 32768, N).astype(np.int16),
        "r": rng.integers(0, 1 << 32, N, dtype=np.uint64).astype(np.uint32),
        "s": rng.normal(0, 1000, N).astype(np.float32),
        "t": rng.integers(1, 256, N).astype(np.uint8),          # divisor column: never zero
        "u": rng.integers(-(1 << 20), 1 << 20, N, dtype=np.int64),
    }
    cols = [ctx.from_numpy(tbl[k]) for k in NAMES2]
    yield ctx, tbl, cols
    ctx.close()


def gen_num2(rng, depth):
    if depth == 0 or rng.random() < 0.3:
        if rng.random() < 0.75:
            return f"(col {rng.choice(NAMES2)})"
        kind = rng.choice(["u8", "i8", "i16", "u32", "f32", "i64"])
        lit = {"u8": rng.randint(1, 255), "i8": rng.randint(-100, 100), "i16": rng.randint(-3000, 3000), "u32": rng.randint(1, 10**9),
               "f32": round(rng.uniform(-50, 50), 1), "i64": rng.randint(-10**6, 10**6)}[kind]
        return f"({kind} {lit})"
    op = rng.choice(["+", "-", "*", "/"])
    l = gen_num2(rng, depth - 1)
    r = rng.choice(["(col t)", f"(u8 {rng.randint(1, 200)})"]) if op == "/" else gen_num2(rng, depth - 1)
    if "(col" not in l + r:
        l = f"(col {rng.choice(NAMES2)})"
    return f"({op} {l} {r})"


@pytest.mark.parametrize("seed", range(200, 216))
def test_random_trees_over_narrow_and_float32_lanes(env2, seed):
    ctx, tbl, cols = env2
    rng = random.Random(seed)
    cmp_ = lambda: f"({rng.choice(['=', '<', '<=', '>', '>='])} {gen_num2(rng, 1)} {gen_num2(rng, 1)})"
    exprs = [gen_num2(rng, 3), gen_num2(rng, 2), cmp_()]
    pred = cmp_() if seed % 2 else None
    table = {k: o.from_numpy(v) for k, v in tbl.items()}
    want, err = oracle_or_error(lambda: o.run_query(exprs, table=table, predicate=pred, worker_threads=1, tail_quirk=False))
    try:
        pipe = ctx.pipe(exprs, columns=NAMES2, dtypes=[COLS2[k][0] for k in NAMES2], predicate=pred)
        outs = [ctx.column(pipe.expr_dtype(i), N) for i in range(len(exprs))]
        ov = [ctx.column(cabi.BOOL, N) if pipe.expr_nullable(i) else None for i in range(len(exprs))]
        pipe.launch_project(cabi.make_source(cols, N), outs, N, out_valid=ov)
        sel, written = pipe.fetch_project()
    except cabi.FuseGpuError as e:
        assert err is not None and str(e) == err, (exprs, pred, str(e), err)
        return
    assert err is None, (exprs, pred, err)
    assert written == sel == want.n_rows
    for i, c in enumerate(want.columns):
        got = outs[i].to_numpy(written)
        assert pipe.expr_dtype(i) == c.dtype, (exprs[i], pipe.expr_dtype(i), c.dtype)
        exp = c.values
        if c.valid is not None or ov[i] is not None:
            wv = np.ones(written, np.uint8) if c.valid is None else c.valid
            gv = np.ones(written, np.uint8) if ov[i] is None else ov[i].to_numpy(written)
            assert np.array_equal(gv, wv), f"validity of {exprs[i]}"
            got, exp = got[wv.astype(bool)], exp[wv.astype(bool)]
        if c.dtype in (o.F32, o.F64):
            assert np.array_equal(got, exp, equal_nan=True), exprs[i]     # same IEEE operations, no FMA contraction: bit-exact
        else:
            assert np.array_equal(got.astype(exp.dtype), exp), exprs[i]
    # the same expressions as aggregate arguments
    aggs = [f"(max {exprs[0]})", f"(min {exprs[1]})", f"(count {exprs[2]})"]
    wa = []
    for e in aggs:
        r, er = oracle_or_error(lambda: o.run_query([e], table=table, predicate=pred, is_aggregate=True, worker_threads=1, tail_quirk=False,
                                                    block_size=1 << 30))
        wa.append(None if er else r.columns[0].to_list()[0])
    pipe = ctx.pipe(aggs, columns=NAMES2, dtypes=[COLS2[k][0] for k in NAMES2], predicate=pred, aggregate=True)
    pipe.launch_aggregate(cabi.make_source(cols, N))
    states, rows = pipe.fetch_aggregate()
    for (dtype, val), w, e in zip(states, wa, aggs):
        if isinstance(w, float) and w != w:
            assert val != val, e
        else:
            assert val == w, (e, val, w)
