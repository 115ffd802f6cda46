"""Parity of the CUDA path (through the C ABI, libfuse_gpu.so) against the CPU oracle.

Bit-exact for every integer / index result.  Floating-point sums are compared with a stated
tolerance (the reduction order differs from the reference's sequential fold); float min/max and
element-wise float arithmetic are exact.
"""
import os

import numpy as np
import pytest

from fuse_query_b200 import cabi
from oracle import binding as o

pytestmark = pytest.mark.gpu

NUM = "(col number)"
README_AGGS = {
    "sum": [f"(sum {NUM})"],
    "max": [f"(max {NUM})"],
    "max_plus_1": [f"(max (+ {NUM} (u64 1)))"],
    "count": [f"(count {NUM})"],
    "avg": [f"(/ (sum {NUM}) (count {NUM}))"],
    "headline": [f"(/ (sum {NUM}) (count {NUM}))", f"(max {NUM})", f"(min {NUM})"],
    "cfg2": [f"(max (+ {NUM} (u64 1)))", f"(min {NUM})", f"(count {NUM})"],
}
README_PRED = f"(< (+ (+ (+ {NUM} (u64 1)) (/ {NUM} (u64 2))) (u64 1)) (u64 100))"
README_PROJ = [f"(alias c1 (+ {NUM} (u64 1)))", f"(alias c2 (/ {NUM} (u64 2)))"]


@pytest.fixture(scope="module")
def ctx():
    c = cabi.Context(0)  # raises (never skips) when the CUDA path is unavailable
    yield c
    c.close()


def splitmix64(n, seed=0x5EED):
    """Non-monotone full-range u64 test column."""
    x = (np.arange(n, dtype=np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return x


def leaf_sexprs(exprs):
    """Aggregator leaves of the select expressions, in the order the device reports them (node order =
    post-order of each expression in turn)."""
    leaves = []

    def walk(toks, i):
        assert toks[i] == "("
        head = toks[i + 1]
        start = i
        i += 2
        if head in ("col",) or head in cabi._TY:
            i += 1
        elif head == "alias":
            i += 1
            i = walk(toks, i)
        elif head in cabi.AGG:
            depth, j = 0, start
            while True:
                depth += toks[j] == "("
                depth -= toks[j] == ")"
                j += 1
                if depth == 0:
                    break
            leaves.append(" ".join(toks[start:j]).replace("( ", "(").replace(" )", ")"))
            return j
        else:
            i = walk(toks, i)
            i = walk(toks, i)
        assert toks[i] == ")"
        return i + 1

    for e in exprs:
        walk(cabi._tokens(e), 0)
    return leaves


def oracle_leaf_values(exprs, **kw):
    # the C ABI scans exactly the rows it is given; the reference's NumbersStream tail quirk
    # is a property of the source and is mirrored by the host-side NumbersStream, not by the kernels
    kw.setdefault("tail_quirk", False)
    out = []
    for leaf in leaf_sexprs(exprs):
        r = o.run_query([leaf], is_aggregate=True, **kw)
        out.append(r.columns[0].to_list()[0])
    return out


def gpu_leaf_values(pipe, src, **kw):
    pipe.launch_aggregate(src, **kw)
    states, rows = pipe.fetch_aggregate()
    return [None if s is None else s[1] for s in states], rows


# ---------------------------------------------------------------------------------------------
# source
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("begin,n", [(0, 0), (0, 1), (7, 2), (5, 3), (10**12, 10001), (123, 1 << 20), (0, 1000003)])
def test_numbers_fill(ctx, begin, n):
    col = ctx.numbers(begin, n)
    got = col.to_numpy()
    assert np.array_equal(got, np.arange(begin, begin + n, dtype=np.uint64))
    col.free()


def test_numbers_fill_unaligned_offset(ctx):
    col = ctx.column(cabi.U64, 1001)
    ctx.check(cabi.lib().fq_numbers_fill(ctx._h, col._h, 1, 50, 1000, None))
    assert np.array_equal(col.to_numpy()[1:], np.arange(50, 1050, dtype=np.uint64))


# ---------------------------------------------------------------------------------------------
# aggregates: the README queries (BASELINE configs[0], [1]) against the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("generated", [False, True], ids=["materialised", "generated"])
@pytest.mark.parametrize("name", list(README_AGGS))
def test_readme_aggregates_10m(ctx, name, generated):
    n = 10_000_000
    exprs = README_AGGS[name]
    pipe = ctx.pipe(exprs, aggregate=True, generated=generated)
    assert pipe.precompiled, "README shapes must come from the precompiled table"
    col = None if generated else ctx.numbers(0, n)
    src = cabi.make_source([] if generated else [col], n, generated=generated, begin=0)
    got, rows = gpu_leaf_values(pipe, src)
    assert rows == n
    assert got == oracle_leaf_values(exprs, total=n, worker_threads=8, use_threads=True)
    pipe.destroy()
    if col:
        col.free()


def test_config0_sum_10m_known_answer(ctx):
    n = 10_000_000
    pipe = ctx.pipe(README_AGGS["sum"], aggregate=True)
    col = ctx.numbers(0, n)
    got, _ = gpu_leaf_values(pipe, cabi.make_source([col], n))
    assert got == [49999995000000]


@pytest.mark.parametrize("n", [0, 1, 2, 3, 511, 2048, 2049, 4097, 65537, 1_000_003])
def test_aggregate_ragged_sizes(ctx, n):
    exprs = README_AGGS["headline"] + [f"(count {NUM})"]
    data = splitmix64(n)
    col = ctx.from_numpy(data)
    pipe = ctx.pipe(exprs, aggregate=True)
    got, rows = gpu_leaf_values(pipe, cabi.make_source([col], n))
    assert rows == n
    if n == 0:
        # arrow sum/min/max over an empty array is None; count is 0 (data_array_aggregate.rs:103-114)
        assert got == [None, 0, None, None, 0]
    else:
        s = int(data.sum(dtype=np.uint64))
        assert got == [s, n, int(data.max()), int(data.min()), n]
        assert got == oracle_leaf_values(exprs, table={"number": o.from_numpy(data)}, worker_threads=8)


def test_state_is_null_before_any_launch(ctx):
    pipe = ctx.pipe(README_AGGS["headline"], aggregate=True)
    states, rows = pipe.fetch_aggregate()
    assert states == [None, None, None, None] and rows == 0  # AggregatorFunction state starts Null


def test_accumulate_across_blocks_equals_one_launch(ctx):
    """Successive reference blocks fold into one running state (function_aggregator.rs:57-100)."""
    n = 300_007
    data = splitmix64(n, seed=7)
    col = ctx.from_numpy(data)
    exprs = [f"(sum (* {NUM} (u64 3)))", f"(min {NUM})", f"(max (/ {NUM} (u64 7)))", f"(count {NUM})"]
    pipe = ctx.pipe(exprs, aggregate=True)
    one, _ = gpu_leaf_values(pipe, cabi.make_source([col], n))
    bounds = [0, 10_000, 10_001, 150_000, n]
    for k in range(len(bounds) - 1):
        part = col.slice(bounds[k], bounds[k + 1] - bounds[k]) if bounds[k] % 2 == 0 else None
        if part is None:  # odd offsets break 16-byte alignment: re-upload that block
            part = ctx.from_numpy(data[bounds[k]:bounds[k + 1]])
        pipe.launch_aggregate(cabi.make_source([part], bounds[k + 1] - bounds[k]), accumulate=k > 0)
    states, rows = pipe.fetch_aggregate()
    assert rows == n
    assert [s[1] for s in states] == one


def test_u64_sum_wraps_like_the_reference(ctx):
    """the sum over 10^10 rows exceeds 2^64 and wraps; checked here on a small column of huge values."""
    data = np.full(1000, (1 << 63) + 12345, dtype=np.uint64)
    col = ctx.from_numpy(data)
    pipe = ctx.pipe([f"(sum {NUM})"], aggregate=True)
    got, _ = gpu_leaf_values(pipe, cabi.make_source([col], len(data)))
    assert got == [(1000 * ((1 << 63) + 12345)) % (1 << 64)]
    assert got == oracle_leaf_values([f"(sum {NUM})"], table={"number": o.from_numpy(data)})


# ---------------------------------------------------------------------------------------------
# expression trees through NVRTC: arithmetic / comparison / logic over random data
# ---------------------------------------------------------------------------------------------
JIT_AGG_CASES = [
    [f"(sum (+ (* {NUM} {NUM}) (u64 17)))", f"(max (- {NUM} (u64 5)))"],
    [f"(min (/ (u64 1000000007) (+ (/ {NUM} (u64 4611686018427387904)) (u64 1))))"],
    [f"(sum (/ {NUM} (u64 3)))", f"(count (+ {NUM} (u64 1)))", f"(max (* (/ {NUM} (u64 1000)) (u64 999)))"],
]


@pytest.mark.parametrize("exprs", JIT_AGG_CASES)
def test_jit_aggregates_match_oracle(ctx, exprs):
    n = 200_003
    data = splitmix64(n, seed=11)
    col = ctx.from_numpy(data)
    pipe = ctx.pipe(exprs, aggregate=True)
    assert not pipe.precompiled
    got, _ = gpu_leaf_values(pipe, cabi.make_source([col], n))
    assert got == oracle_leaf_values(exprs, table={"number": o.from_numpy(data)}, worker_threads=8)


def test_filtered_aggregate_matches_oracle(ctx):
    n = 123_457
    data = splitmix64(n, seed=3) >> np.uint64(40)
    col = ctx.from_numpy(data)
    pred = f"(and (> {NUM} (u64 1000)) (<= (/ {NUM} (u64 2)) (u64 4000000)))"
    exprs = [f"(min {NUM})", f"(max {NUM})", f"(count {NUM})"]
    pipe = ctx.pipe(exprs, predicate=pred, aggregate=True)
    got, rows = gpu_leaf_values(pipe, cabi.make_source([col], n))
    mask = (data > 1000) & ((data // 2) <= 4000000)
    assert rows == int(mask.sum())
    assert got == [int(data[mask].min()), int(data[mask].max()), rows]
    # min/max/count survive fully-filtered blocks in the reference (SURVEY F8 only poisons Sum)
    assert got == oracle_leaf_values(exprs, table={"number": o.from_numpy(data)}, predicate=pred, worker_threads=8)


def test_divide_by_zero_is_an_error(ctx):
    data = np.array([5, 4, 0, 2], dtype=np.uint64)
    col = ctx.from_numpy(data)
    pipe = ctx.pipe([f"(sum (/ (u64 100) {NUM}))"], aggregate=True)
    pipe.launch_aggregate(cabi.make_source([col], 4))
    with pytest.raises(cabi.FuseGpuError) as ei:
        pipe.fetch_aggregate()
    assert ei.value.status == cabi.ERR_DIVIDE_BY_ZERO
    with pytest.raises(o.OracleError) as oi:
        o.run_query([f"(sum (/ (u64 100) {NUM}))"], table={"number": o.from_numpy(data)}, is_aggregate=True)
    assert str(ei.value) == str(oi.value) == "Internal Error: Divide by zero error"


def test_type_errors_reproduce_reference_text(ctx):
    with pytest.raises(cabi.FuseGpuError) as ei:
        ctx.pipe([f"(sum (and {NUM} {NUM}))"], aggregate=True)
    with pytest.raises(o.OracleError) as oi:
        o.run_query([f"(sum (and {NUM} {NUM}))"], total=8, is_aggregate=True)
    assert str(ei.value) == str(oi.value)
    with pytest.raises(cabi.FuseGpuError) as ei:
        ctx.pipe([NUM], predicate=f"(+ {NUM} (u64 1))")
    assert str(ei.value) == "Internal Error: cannot downcast to boolean array"


MIXED = {
    "a": (cabi.I64, np.int64), "b": (cabi.I32, np.int32), "c": (cabi.F64, np.float64), "d": (cabi.U16, np.uint16),
}


def mixed_table(n, seed=5):
    rng = np.random.default_rng(seed)
    return {
        "a": rng.integers(-10**12, 10**12, n, dtype=np.int64),
        "b": rng.integers(-50000, 50000, n, dtype=np.int32),
        "c": rng.normal(0, 1000, n),
        "d": rng.integers(1, 60000, n, dtype=np.uint16),
    }


def test_mixed_type_columns_and_coercion(ctx):
    """Int64/Int32/Float64/UInt16 columns: coercion lattice of data_type.rs:27-87, integer lanes exact,
    float sum within 1e-9 relative (reduction order)."""
    n = 100_003
    tbl = mixed_table(n)
    names = list(tbl)
    cols = [ctx.from_numpy(tbl[k]) for k in names]
    dtypes = [MIXED[k][0] for k in names]
    exprs = ["(sum (+ (col a) (col b)))", "(min (* (col b) (col d)))", "(max (- (col a) (col d)))", "(sum (col c))",
             "(max (+ (col c) (col b)))", "(min (/ (col a) (col d)))", "(count (col d))"]
    pred = "(< (col b) (col d))"
    pipe = ctx.pipe(exprs, columns=names, dtypes=dtypes, predicate=pred, aggregate=True)
    got, rows = gpu_leaf_values(pipe, cabi.make_source(cols, n))
    want = oracle_leaf_values(exprs, table={k: o.from_numpy(v) for k, v in tbl.items()}, predicate=pred, worker_threads=1)
    assert [pipe.expr_dtype(i) for i in range(len(exprs))] == [cabi.I64, cabi.I32, cabi.I64, cabi.F64, cabi.F64, cabi.I64, cabi.U64]
    for i, (g, w) in enumerate(zip(got, want)):
        if i == 3:
            assert abs(g - w) <= 1e-9 * max(1.0, abs(w))  # float sum: stated tolerance
        else:
            assert g == w, (i, g, w)


# ---------------------------------------------------------------------------------------------
# filter -> projection -> limit
# ---------------------------------------------------------------------------------------------
def run_project(ctx, pipe, src, n_exprs, capacity, **kw):
    outs = [ctx.column(pipe.expr_dtype(i), max(capacity, 1)) for i in range(n_exprs)]
    pipe.launch_project(src, outs, capacity, **kw)
    sel, written = pipe.fetch_project()
    res = [c.to_numpy(written) for c in outs]
    for c in outs:
        c.free()
    return sel, written, res


@pytest.mark.parametrize("generated", [False, True], ids=["materialised", "generated"])
def test_readme_filter_projection_limit(ctx, generated):
    """BASELINE configs[2] at the README's own size (10^7 rows): rows (1,0),(2,0),(3,1); 66 rows match."""
    n = 10_000_000
    pipe = ctx.pipe(README_PROJ, predicate=README_PRED, generated=generated)
    assert pipe.precompiled
    col = None if generated else ctx.numbers(0, n)
    src = cabi.make_source([] if generated else [col], n, generated=generated)
    sel, written, (c1, c2) = run_project(ctx, pipe, src, 2, capacity=3, limit=3)
    assert (sel, written) == (66, 3)
    assert list(zip(c1.tolist(), c2.tolist())) == [(1, 0), (2, 0), (3, 1)]
    want = o.run_query(README_PROJ, total=n, predicate=README_PRED, limit=3, worker_threads=8)
    assert list(zip(c1.tolist(), c2.tolist())) == want.rows() and want.names == ["c1", "c2"]
    # no limit: all 66 rows, in order
    sel, written, (c1, c2) = run_project(ctx, pipe, src, 2, capacity=1000)
    want = o.run_query(README_PROJ, total=n, predicate=README_PRED, worker_threads=8)
    assert sel == written == 66 and list(zip(c1.tolist(), c2.tolist())) == want.rows()
    # early exit gives the same rows
    sel, written, (c1, c2) = run_project(ctx, pipe, src, 2, capacity=3, limit=3, early_exit=True)
    assert written == 3 and sel >= 3 and list(zip(c1.tolist(), c2.tolist())) == [(1, 0), (2, 0), (3, 1)]


@pytest.mark.parametrize("n", [0, 1, 2, 3, 2047, 2048, 2049, 6145, 250_001])
@pytest.mark.parametrize("limit", [-1, 0, 1, 1000])
def test_compaction_is_order_preserving(ctx, n, limit):
    """Dense, irregular selection (multiples of 3 of a shuffled column): arrow filter keeps row order."""
    data = splitmix64(n, seed=21) >> np.uint64(8)
    col = ctx.from_numpy(data)
    pred = f"(= (* (/ {NUM} (u64 3)) (u64 3)) {NUM})"
    exprs = [NUM, f"(+ (/ {NUM} (u64 3)) (u64 1))", f"(< {NUM} (u64 36028797018963968))"]
    pipe = ctx.pipe(exprs, predicate=pred)
    cap = max(n, 1)
    sel, written, res = run_project(ctx, pipe, cabi.make_source([col], n), 3, capacity=cap, limit=limit)
    mask = (data % 3) == 0
    keep = data[mask]
    assert sel == len(keep)
    if limit >= 0:
        keep = keep[:limit]
    assert written == len(keep)
    assert np.array_equal(res[0], keep)
    assert np.array_equal(res[1], keep // 3 + 1)
    assert np.array_equal(res[2].astype(bool), keep < (1 << 55))
    want = o.run_query(exprs, table={"number": o.from_numpy(data)}, predicate=pred, limit=None if limit < 0 else limit,
                       worker_threads=1, tail_quirk=False)
    assert want.n_rows == written
    assert np.array_equal(want.columns[0].values, res[0]) and np.array_equal(want.columns[1].values, res[1])


def test_projection_without_filter(ctx):
    n = 70_001
    data = splitmix64(n, seed=2)
    col = ctx.from_numpy(data)
    exprs = [f"(+ {NUM} (u64 1))", f"(/ {NUM} (u64 2))", f"(>= {NUM} (u64 9223372036854775808))"]
    pipe = ctx.pipe(exprs)
    sel, written, res = run_project(ctx, pipe, cabi.make_source([col], n), 3, capacity=n)
    assert sel == written == n
    assert np.array_equal(res[0], data + np.uint64(1)) and np.array_equal(res[1], data // 2)
    assert np.array_equal(res[2].astype(bool), data >= (1 << 63))
    want = o.run_query(exprs, table={"number": o.from_numpy(data)}, worker_threads=1, tail_quirk=False)
    assert np.array_equal(want.columns[0].values, res[0]) and np.array_equal(want.columns[1].values, res[1])
    assert np.array_equal(want.columns[2].values, res[2])


def test_capacity_truncates_but_counts_everything(ctx):
    n = 50_000
    col = ctx.numbers(0, n)
    pipe = ctx.pipe([NUM], predicate=f"(> {NUM} (u64 9))")
    sel, written, res = run_project(ctx, pipe, cabi.make_source([col], n), 1, capacity=100)
    assert sel == n - 10 and written == 100
    assert np.array_equal(res[0], np.arange(10, 110, dtype=np.uint64))


# ---------------------------------------------------------------------------------------------
# BASELINE sizes: closed-form known answers + size-independent properties
# ---------------------------------------------------------------------------------------------
def closed_form(n):
    s = (n * (n - 1) // 2) % (1 << 64)
    return {"sum": s, "count": n, "max": n - 1, "min": 0}


@pytest.mark.parametrize("n", [10**9, 10**10], ids=["1e9", "1e10"])
def test_baseline_sizes_generated(ctx, n):
    exprs = README_AGGS["headline"] + README_AGGS["cfg2"]
    pipe = ctx.pipe(README_AGGS["headline"], aggregate=True, generated=True)
    got, rows = gpu_leaf_values(pipe, cabi.make_source([], n, generated=True))
    cf = closed_form(n)
    assert rows == n and got == [cf["sum"], n, cf["max"], 0]
    pipe2 = ctx.pipe(README_AGGS["cfg2"], aggregate=True, generated=True)
    got, _ = gpu_leaf_values(pipe2, cabi.make_source([], n, generated=True))
    assert got == [n, 0, n]
    if n == 10**10:  # the u64 sum wraps, sum/count is integer division
        assert cf["sum"] == 13106511847580896768 and cf["sum"] // n == 1310651184


def test_baseline_1e9_materialised_and_split_invariance(ctx):
    """configs[1]/[2] at full size on a materialised shard; the aggregate of the whole equals the fold of
    8 partition launches (merge is associative + commutative on wrapping u64)."""
    n = 10**9
    col = ctx.numbers(0, n)
    pipe = ctx.pipe(README_AGGS["headline"], aggregate=True)
    whole, rows = gpu_leaf_values(pipe, cabi.make_source([col], n))
    cf = closed_form(n)
    assert rows == n and whole == [cf["sum"], n, cf["max"], 0]
    parts = o.generate_parts(n)
    for k, (b, e) in enumerate(parts):
        pipe.launch_aggregate(cabi.make_source([col.slice(b, e - b + 1)], e - b + 1), accumulate=k > 0)
    states, rows = pipe.fetch_aggregate()
    assert rows == n and [s[1] for s in states] == whole
    # configs[2]: filter + projection + limit over 10^9 rows
    p3 = ctx.pipe(README_PROJ, predicate=README_PRED)
    sel, written, (c1, c2) = run_project(ctx, p3, cabi.make_source([col], n), 2, capacity=3, limit=3)
    assert sel == 66 and list(zip(c1.tolist(), c2.tolist())) == [(1, 0), (2, 0), (3, 1)]
    col.free()


def test_baseline_1e10_materialised_headline(ctx):
    """configs[3]: 80 GB shard resident in HBM; sum wraps to 13106511847580896768."""
    n = 10**10
    col = ctx.numbers(0, n)
    pipe = ctx.pipe(README_AGGS["headline"], aggregate=True)
    got, rows = gpu_leaf_values(pipe, cabi.make_source([col], n))
    assert rows == n and got == [13106511847580896768, n, n - 1, 0]
    assert got[0] // got[1] == 1310651184
    # filter + projection over the same shard: row indexes beyond 2^32 through segment claims, look-back and scatter
    k = 1_000_000_007
    for generated in (False, True):
        src = cabi.make_source([] if generated else [col], n, generated=generated)
        p = ctx.pipe([NUM, f"(+ {NUM} (u64 1))"], predicate=f"(= (* (/ {NUM} (u64 {k})) (u64 {k})) {NUM})", generated=generated)
        sel, written, (c0, c1) = run_project(ctx, p, src, 2, capacity=16)
        want = np.arange(0, n, k, dtype=np.uint64)
        assert sel == written == len(want) == 10 and np.array_equal(c0, want) and np.array_equal(c1, want + 1)
        p = ctx.pipe([NUM], predicate=f"(>= {NUM} (u64 {n - 7}))", generated=generated)
        sel, written, (c0,) = run_project(ctx, p, src, 1, capacity=16)
        assert sel == written == 7 and np.array_equal(c0, np.arange(n - 7, n, dtype=np.uint64))
    col.free()


# ---------------------------------------------------------------------------------------------
# SURVEY §8b "Threading": one context, many host threads (the reference runs one tokio task per pipe,
# processor_merge.rs:46-62).  Compiles (NVRTC included), launches on per-thread streams and fetches run concurrently.
# ---------------------------------------------------------------------------------------------
def test_context_is_usable_from_many_host_threads(ctx):
    import threading
    import torch
    n_threads, n = 8, 3_000_017
    shards = [(t * n, n) for t in range(n_threads)]
    cols = [ctx.numbers(b, m) for b, m in shards]
    ctx.synchronize()
    results, errors = [None] * n_threads, []

    def work(t):
        try:
            stream = torch.cuda.Stream()
            begin, m = shards[t]
            # a different tree per thread: some precompiled, some through NVRTC, all at once
            exprs = [f"(sum (+ {NUM} (u64 {t})))", f"(max {NUM})", f"(min (* {NUM} (u64 {t + 1})))", f"(count {NUM})"]
            pipe = ctx.pipe(exprs, aggregate=True)
            sel = ctx.pipe([f"(+ {NUM} (u64 {t}))"], predicate=f"(= (* (/ {NUM} (u64 1000)) (u64 1000)) {NUM})")
            out = ctx.column(cabi.U64, m // 1000 + 2)
            src = cabi.make_source([cols[t]], m)
            for _ in range(5):
                pipe.launch_aggregate(src, stream=stream.cuda_stream)
                sel.launch_project(src, [out], m // 1000 + 2, stream=stream.cuda_stream)
                states, rows = pipe.fetch_aggregate()
                k, w = sel.fetch_project()
            results[t] = ([v for _, v in states], rows, k, out.to_numpy(w, stream=stream.cuda_stream))
        except Exception as e:   # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for t, (begin, m) in enumerate(shards):
        x = np.arange(begin, begin + m, dtype=np.uint64)
        vals, rows, k, got = results[t]
        assert rows == m
        assert vals == [int((x + np.uint64(t)).sum(dtype=np.uint64)), int(x.max()), int((x * np.uint64(t + 1)).min()), m]
        want = x[x % 1000 == 0] + np.uint64(t)
        assert k == len(want) and np.array_equal(got, want)


def _streams(k):
    import torch
    return [torch.cuda.Stream(device=0) for _ in range(k)]


def _local_groups(ctx, world, row_bytes=1 << 16):
    """`world` ranks of one group living in this process on one GPU (each rank's kernels go to its own stream): the exchange
    windows are plain device addresses instead of IPC mappings, the kernels are those of the multi-process path."""
    groups = [ctx.group(r, world, row_bytes) for r in range(world)]
    wins = [g.window for g in groups]
    for g in groups:
        g.connect_ptrs(wins)
    return groups


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("variant", ["tma", "u4"])
def test_aggregate_kernels_merge_their_states_across_a_group(ctx, world, variant):
    """fq_pipe_set_group: every aggregate launch ends with the merge point inside the kernel (exchange of the running
    states over peer memory + final fold in rank order, processor_merge.rs:37-66 -> transform_aggregate_final.rs:50-78).
    Every rank ends up with the merged state of all ranks; several operations in a row reuse the two row parities."""
    import torch
    n_total = 8 * 625_003 + 5
    groups = _local_groups(ctx, world)
    streams = _streams(world)
    bounds = [n_total * r // world for r in range(world + 1)]
    bounds[1:-1] = [b // 2 * 2 for b in bounds[1:-1]]
    cols = [ctx.numbers(11 + bounds[r], bounds[r + 1] - bounds[r]) for r in range(world)]
    pipes = []
    for r in range(world):
        p = ctx.pipe(README_AGGS["headline"] + [f"(count {NUM})", f"(max (+ {NUM} (u64 1)))"], aggregate=True)
        p.set_variant(variant)
        p.set_group(groups[r])
        pipes.append(p)
    x = np.arange(11, 11 + n_total, dtype=np.uint64)
    want = [int(x.sum(dtype=np.uint64)), n_total, int(x.max()), int(x.min()), n_total, int(x.max()) + 1]
    for op in range(5):   # epochs 1..5: both parities, rows reused
        for r in range(world):
            pipes[r].launch_aggregate(cabi.make_source([cols[r]], bounds[r + 1] - bounds[r]), stream=streams[r].cuda_stream)
        for r in range(world):
            states, rows = pipes[r].fetch_merged()
            assert rows == n_total and [s[1] for s in states] == want, (op, r)
            local, lrows = pipes[r].fetch_aggregate()
            assert lrows == bounds[r + 1] - bounds[r]
    # a filtered Sum whose predicate keeps nothing on some ranks: Type(None) only if NO rank saw a row
    pipes2 = []
    for r in range(world):
        p = ctx.pipe([f"(sum {NUM})", f"(min {NUM})"], predicate=f"(< {NUM} (u64 100))", aggregate=True)
        p.set_group(groups[r])
        pipes2.append(p)
    for r in range(world):
        pipes2[r].launch_aggregate(cabi.make_source([cols[r]], bounds[r + 1] - bounds[r]), stream=streams[r].cuda_stream)
    for r in range(world):
        states, rows = pipes2[r].fetch_merged()
        assert rows == 89 and [s[1] for s in states] == [sum(range(11, 100)), 11]
    torch.cuda.synchronize()
    for p in pipes + pipes2:
        p.destroy()
    for g in groups:
        g.destroy()


def test_group_merge_reports_a_missing_rank_instead_of_hanging(ctx):
    """A rank that never launches makes the others fail after FQ_GROUP_TIMEOUT_MS — never a hung GPU."""
    os.environ["FQ_GROUP_TIMEOUT_MS"] = "200"
    try:
        groups = _local_groups(ctx, 2)
    finally:
        os.environ.pop("FQ_GROUP_TIMEOUT_MS", None)
    col = ctx.numbers(0, 100_000)
    p = ctx.pipe(README_AGGS["sum"], aggregate=True)
    p.set_group(groups[0])
    p.launch_aggregate(cabi.make_source([col], 100_000))
    with pytest.raises(cabi.FuseGpuError) as e:
        p.fetch_merged()
    assert "merge timed out" in str(e.value)
    states, rows = p.fetch_aggregate()      # the local state is intact
    assert rows == 100_000 and states[0][1] == 100_000 * 99_999 // 2
    p.destroy()
    for g in groups:
        g.destroy()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("limit", [3, 70, 1000, -1])
def test_group_gathers_filtered_rows_in_rank_order(ctx, world, limit):
    """fq_group_gather_project: MergeProcessor + the LimitTransform after it for projection pipes — every rank's kept rows,
    ranks in partition order, cut at LIMIT (pipeline_builder.rs:31-41).  Checked against the oracle's single-pipe run."""
    import torch
    n_total = 1_600_000
    per = n_total // world
    groups = _local_groups(ctx, world)
    streams = _streams(world)
    pred = f"(= (* (/ {NUM} (u64 25000)) (u64 25000)) {NUM})"      # every 25 000th number: 64 rows, 64 / world per rank
    exprs = [f"(alias c1 (+ {NUM} (u64 1)))", f"(alias small (< {NUM} (u64 800000)))"]
    cap = 1000 if limit < 0 else limit
    finals = []
    for r in range(world):
        col = ctx.numbers(r * per, per)
        p = ctx.pipe(exprs, predicate=pred)
        outs = [ctx.column(p.expr_dtype(i), cap) for i in range(2)]
        fin = [ctx.column(p.expr_dtype(i), cap * world) for i in range(2)]
        p.launch_project(cabi.make_source([col], per), outs, cap, limit=limit, stream=streams[r].cuda_stream)
        groups[r].gather_project(p, outs, fin, limit=limit, stream=streams[r].cuda_stream)
        finals.append(fin)
    want = o.run_query(exprs, total=n_total, predicate=pred, limit=None if limit < 0 else limit, worker_threads=1, tail_quirk=False).rows()
    for r in range(world):
        sel, nfin = groups[r].fetch_gather()
        assert nfin == len(want)
        if limit < 0 or limit >= 64:
            assert sel == 64
        got = list(zip(finals[r][0].to_numpy(nfin).tolist(), [bool(b) for b in finals[r][1].to_numpy(nfin)]))
        assert got == [(a, bool(b)) for a, b in want]
    torch.cuda.synchronize()
    for g in groups:
        g.destroy()


def test_jit_cubins_are_cached_on_disk(tmp_path):
    """A specialisation outside the precompiled table is built by NVRTC once and then loaded from
    $FQ_JIT_CACHE_DIR by every later context / process; FQ_JIT_CACHE=0 switches the cache off.  Same results either way."""
    exprs = [f"(sum (* {NUM} (u64 31337)))", f"(max (+ {NUM} (u64 424242)))"]      # not in aot_pipes.txt
    n = 1_000_003
    x = np.arange(n, dtype=np.uint64)
    want = [int((x * np.uint64(31337)).sum(dtype=np.uint64)), int(x.max()) + 424242]
    old = {k: os.environ.get(k) for k in ("FQ_JIT_CACHE_DIR", "FQ_JIT_CACHE")}
    os.environ["FQ_JIT_CACHE_DIR"] = str(tmp_path / "jit")
    os.environ.pop("FQ_JIT_CACHE", None)
    try:
        kinds = []
        for _ in range(2):                      # a context caches modules in memory, so use a fresh one each time
            c = cabi.Context(0)
            pipe = c.pipe(exprs, aggregate=True, generated=True)
            pipe.launch_aggregate(cabi.make_source([], n, generated=True))
            states, rows = pipe.fetch_aggregate()
            assert [s[1] for s in states] == want and rows == n
            kinds.append(pipe.build_kind)
            c.close()
        assert kinds == [1, 2]
        assert len(list((tmp_path / "jit").glob("*.cubin"))) == 1
        os.environ["FQ_JIT_CACHE"] = "0"
        c = cabi.Context(0)
        assert c.pipe(exprs, aggregate=True, generated=True).build_kind == 1
        assert c.pipe(README_AGGS["headline"], aggregate=True).build_kind == 0
        c.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
