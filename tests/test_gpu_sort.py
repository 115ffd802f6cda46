"""fq_sort_indices / fq_column_take (ORDER BY on the device) against oracle/sort.py: every key type, NULLs as bytes and as
bitmaps, descending, several keys, stability, sizes around the tile (4096) and scan-tile (8192 counters) boundaries.
Integer work: the permutation must be identical, not merely a valid sort."""
import numpy as np
import pytest

from fuse_query_b200 import cabi
from oracle.sort import sort_indices

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = cabi.Context(0)
    yield c
    c.close()


def gpu_perm(ctx, keys, n, desc=None):
    idx = ctx.sort_indices(keys, n, desc)
    out = idx.to_numpy(n)
    idx.free()
    return out.astype(np.int64)


@pytest.mark.parametrize("n", [0, 1, 31, 4095, 4096, 4097, 70_001, 1_300_003])
def test_one_unsigned_key_is_sorted_stably(ctx, n):
    rng = np.random.default_rng(n)
    a = rng.integers(0, 1000, n).astype(np.uint64)          # many ties: stability decides the permutation
    col = ctx.from_numpy(a)
    assert gpu_perm(ctx, [col], n).tolist() == sort_indices([a]).tolist()
    assert gpu_perm(ctx, [col], n, [True]).tolist() == sort_indices([a], descending=[True]).tolist()
    col.free()


@pytest.mark.parametrize("dtype", [np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.uint64, np.int64, np.float32, np.float64])
def test_every_key_type_with_nulls(ctx, dtype):
    rng = np.random.default_rng(3)
    n = 50_003
    if np.dtype(dtype).kind == "f":
        a = rng.normal(size=n).astype(dtype)
        a[::97] = np.nan
        a[1::97] = -np.inf
        a[2::97] = -0.0
        a[3::97] = 0.0
    else:
        info = np.iinfo(dtype)
        a = rng.integers(info.min, info.max, n, dtype=dtype, endpoint=True)
        a[::50] = info.min
        a[1::50] = info.max
    ok = rng.random(n) > 0.15
    for form in ("bytes", "bits"):
        if form == "bytes":
            col = ctx.from_numpy(a, ok)
        else:
            col = ctx.from_numpy(a, valid_bitmap=np.packbits(ok, bitorder="little"))
        for desc in (False, True):
            want = sort_indices([a], [ok], [desc])
            got = gpu_perm(ctx, [col], n, [desc])
            assert got.tolist() == want.tolist(), (dtype, form, desc)
        rows = ctx.sort_indices([col], n)
        taken = ctx.take(col, rows, n)
        want = sort_indices([a], [ok])
        tv = taken.validity.to_numpy(n).astype(bool)
        assert tv.tolist() == ok[want].tolist()
        got_vals = taken.to_numpy(n)
        assert got_vals[tv].tobytes() == a[want][tv].tobytes()          # bit patterns (NaN payloads, -0.0) travel untouched
        for c in (rows, taken, col):
            c.free()


def test_three_keys_mixed_directions(ctx):
    rng = np.random.default_rng(11)
    n = 200_000
    k0 = rng.integers(0, 5, n).astype(np.uint8)
    k1 = rng.integers(-3, 3, n).astype(np.int16)
    k1_ok = rng.random(n) > 0.3
    k2 = rng.normal(size=n)
    cols = [ctx.from_numpy(k0), ctx.from_numpy(k1, k1_ok), ctx.from_numpy(k2)]
    for desc in ([False, False, False], [True, False, True], [False, True, False]):
        want = sort_indices([k0, k1, k2], [None, k1_ok, None], desc)
        assert gpu_perm(ctx, cols, n, desc).tolist() == want.tolist()
    for c in cols:
        c.free()


@pytest.mark.parametrize("limit", [1, 10, 1000, 40_000])
def test_order_by_limit_is_the_head_of_the_full_order(ctx, limit):
    """fq_sort_indices_limit: radix select (one NOT NULL key, small result) or full sort + cut — either way the first `limit`
    entries of the stable order, ties resolved by input order."""
    rng = np.random.default_rng(limit)
    n = 3_000_017
    cases = {"ties": rng.integers(0, 1000, n).astype(np.uint64),                 # ~3000 rows per value: stability decides
             "u64": rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + np.uint64(1),
             "i32": rng.integers(-2**31, 2**31 - 1, n).astype(np.int32),
             "f64": rng.normal(size=n),
             "same": np.full(n, 7, dtype=np.uint16)}
    for name, a in cases.items():
        col = ctx.from_numpy(a)
        for desc in (False, True):
            want = sort_indices([a], descending=[desc])[:limit]
            idx, count = ctx.sort_indices_limit([col], n, limit, [desc])
            assert count == limit and idx.to_numpy(count).astype(np.int64).tolist() == want.tolist(), (name, desc)
            idx.free()
        col.free()
    # nullable or several keys: the full sort, cut
    a = cases["ties"]
    ok = rng.random(n) > 0.5
    b = cases["i32"]
    cols = [ctx.from_numpy(a, ok), ctx.from_numpy(b)]
    idx, count = ctx.sort_indices_limit(cols, n, limit, [True, False])
    assert idx.to_numpy(count).astype(np.int64).tolist() == sort_indices([a, b], [ok, None], [True, False])[:limit].tolist()
    idx.free()
    idx, count = ctx.sort_indices_limit([cols[1]], 100, 1000)          # limit beyond the rows
    assert count == 100 and idx.to_numpy(100).astype(np.int64).tolist() == sort_indices([b[:100]]).tolist()
    for c in cols + [idx]:
        c.free()


def test_numbers_descending_at_scale_and_errors(ctx):
    n = 30_000_000
    col = ctx.numbers(5, n)
    idx = ctx.sort_indices([col], n, [True])
    head = idx.to_numpy(3)
    assert head.tolist() == [n - 1, n - 2, n - 3]
    taken = ctx.take(col, idx, n)
    t = taken.to_numpy(n)
    assert t[0] == n + 4 and t[-1] == 5 and bool(np.all(t[:-1] > t[1:]))
    with pytest.raises(cabi.FuseGpuError) as e:
        ctx.take(col, col, 10)
    assert "UInt32" in str(e.value)
    for c in (idx, taken, col):
        c.free()
