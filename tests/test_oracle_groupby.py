"""The GROUP BY oracle (oracle/groupby.py) against an independent implementation: pyarrow's hash aggregation.
GROUP BY is planned but not executed by the reference (pipeline_builder.rs:50-65), so no reference vector exists for it;
the oracle applies the reference's aggregate protocol per group, and this file checks that statement on random data
with NULL keys and NULL values."""
import numpy as np
import pyarrow as pa
import pytest

from oracle import binding as o
from oracle.groupby import run_group_by


@pytest.mark.parametrize("seed", range(6))
def test_group_by_oracle_matches_pyarrow(seed):
    rng = np.random.default_rng(seed)
    n = 3000
    k1 = rng.integers(0, 9, n).astype(np.uint16)
    k2 = rng.integers(-3, 3, n).astype(np.int32)
    k1_ok = rng.random(n) > 0.1
    v = rng.integers(-1000, 1000, n).astype(np.int64)
    v_ok = rng.random(n) > (0.3 if seed % 2 else 0.0)
    w = rng.integers(0, 1 << 62, n).astype(np.uint64)
    table = {"k1": o.Array(o.U16, k1, k1_ok.astype(np.uint8)), "k2": o.from_numpy(k2),
             "v": o.Array(o.I64, v, v_ok.astype(np.uint8)), "w": o.from_numpy(w)}
    names, rows = run_group_by(["(col k1)", "(col k2)"], ["(sum (col v))", "(min (col v))", "(max (col v))", "(count (col v))", "(sum (col w))"],
                               table=table, predicate="(< (col k2) (i32 2))")
    assert names == ["k1", "k2", "Sum(v)", "Min(v)", "Max(v)", "Count(v)", "Sum(w)"]
    keep = k2 < 2
    t = pa.table({"k1": pa.array(k1, mask=~k1_ok), "k2": pa.array(k2), "v": pa.array(v, mask=~v_ok), "w": pa.array(w)}).filter(pa.array(keep))
    got = t.group_by(["k1", "k2"]).aggregate([("v", "sum"), ("v", "min"), ("v", "max"), ("v", "count_all" if False else "count"), ("w", "list")])
    # pyarrow: count(v) counts non-null values; the reference's Count is the block length (data_array_aggregate.rs:29) — compare
    # against the group size instead; u64 sums wrap in the reference, so they are folded here from the lists
    sizes = t.group_by(["k1", "k2"]).aggregate([([], "count_all")])
    size_of = {(a, b): c for a, b, c in zip(sizes["k1"].to_pylist(), sizes["k2"].to_pylist(), sizes["count_all"].to_pylist())}
    want = []
    for a, b, s, mn, mx, ws in zip(got["k1"].to_pylist(), got["k2"].to_pylist(), got["v_sum"].to_pylist(), got["v_min"].to_pylist(),
                                   got["v_max"].to_pylist(), got["w_list"].to_pylist()):
        want.append((a, b, s, mn, mx, size_of[(a, b)], sum(ws) % (1 << 64)))
    want.sort(key=lambda r: tuple((x is not None, x) for x in r[:2]))
    assert rows == want


def test_group_by_oracle_outer_arithmetic_and_empty_input():
    names, rows = run_group_by(["(/ (col number) (u64 10))"], ["(/ (sum (col number)) (count (col number)))"], total=35)
    assert names == ["number / 10", "Sum(number) / Count(number)"]
    assert rows == [(0, 4), (1, 14), (2, 24), (3, 32)]
    assert run_group_by(["(col number)"], ["(sum (col number))"], total=10, predicate="(> (col number) (u64 100))")[1] == []
