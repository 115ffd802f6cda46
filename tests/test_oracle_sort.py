"""oracle/sort.py (the stated ORDER BY semantics: the reference does not sort) against pyarrow's sort_indices where both
define the order, and against hand-written cases for what pyarrow orders differently (NaN, -0.0)."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

from oracle.sort import sort_indices


def test_integers_and_nulls_against_pyarrow_single_and_multi_key():
    rng = np.random.default_rng(7)
    n = 5000
    a = rng.integers(-50, 50, n).astype(np.int32)
    b = rng.integers(0, 1 << 40, n).astype(np.uint64)
    a_ok = rng.random(n) > 0.1
    for desc in ([False, False], [True, False], [False, True], [True, True]):
        table = pa.table({"a": pa.array(a, mask=~a_ok), "b": pa.array(b)})
        want = pc.sort_indices(table, sort_keys=[("a", "descending" if desc[0] else "ascending"), ("b", "descending" if desc[1] else "ascending")],
                               null_placement="at_start").to_numpy()
        got = sort_indices([a, b], [a_ok, None], desc)
        assert got.tolist() == want.tolist()


def test_stability_and_float_total_order():
    k = np.array([2, 1, 2, 1, 2], dtype=np.uint8)
    assert sort_indices([k]).tolist() == [1, 3, 0, 2, 4]
    assert sort_indices([k], descending=[True]).tolist() == [0, 2, 4, 1, 3]
    f = np.array([1.5, -0.0, 0.0, np.inf, -np.inf, np.nan, -1.0], dtype=np.float64)
    assert sort_indices([f]).tolist() == [4, 6, 1, 2, 0, 3, 5]
    f32 = f.astype(np.float32)
    assert sort_indices([f32], descending=[True]).tolist() == [5, 3, 0, 2, 1, 6, 4]
    finite = np.random.default_rng(1).normal(size=1000)
    assert sort_indices([finite]).tolist() == np.argsort(finite, kind="stable").tolist()
    # NULLs come first in both directions and keep their input order
    v = np.array([5, 9, 1, 7], dtype=np.int64)
    ok = np.array([True, False, True, False])
    assert sort_indices([v], [ok]).tolist() == [1, 3, 2, 0]
    assert sort_indices([v], [ok], [True]).tolist() == [1, 3, 0, 2]
