"""ORDER BY oracle — TEST INFRASTRUCTURE ONLY (never imported by the product path).

The reference does not sort (README.md:28 "[ ] Sorting"; sqlparser accepts ORDER BY, plan_parser.rs never reads
`query.order_by`), so there is no reference behaviour to restate: this file STATES the semantics the device operator
implements and tests/test_oracle_sort.py cross-checks it against pyarrow's `sort_indices` where both define the order
(integers and finite floats, NULLs placed at the start).  parity unpinned by the reference.

  * order of one key: arrow's sort with the Rust crate's default SortOptions (arrow 2.0 `compute::sort`):
    ascending unless DESC, NULLs first (for both directions);
  * ties keep their input order (stable);
  * floats: IEEE total order (-NaN < -inf < ... < -0.0 < +0.0 < ... < +inf < +NaN); DESC reverses it;
  * several keys: lexicographic, first key most significant.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


def order_code(values: np.ndarray) -> np.ndarray:
    """order-preserving uint64 image of a numeric / bool array"""
    v = np.ascontiguousarray(values)
    if v.dtype == np.bool_:
        return v.astype(np.uint64)
    if v.dtype.kind == "u":
        return v.astype(np.uint64)
    if v.dtype.kind == "i":
        bits = 8 * v.dtype.itemsize
        u = v.view(np.dtype(f"u{v.dtype.itemsize}")).astype(np.uint64)
        return u ^ np.uint64(1 << (bits - 1))
    if v.dtype.kind == "f":
        bits = 8 * v.dtype.itemsize
        u = v.view(np.dtype(f"u{v.dtype.itemsize}")).astype(np.uint64)
        full = np.uint64((1 << bits) - 1)
        sign = np.uint64(1 << (bits - 1))
        neg = (u & sign) != 0
        return np.where(neg, ~u & full, u | sign)
    raise TypeError(f"no order for {v.dtype}")


def sort_indices(keys: Sequence[np.ndarray], valids: Optional[Sequence[Optional[np.ndarray]]] = None,
                 descending: Optional[Sequence[bool]] = None) -> np.ndarray:
    """row indexes in sorted order"""
    n = len(keys[0])
    perm = np.arange(n, dtype=np.int64)
    for j in range(len(keys) - 1, -1, -1):
        code = order_code(np.asarray(keys[j]))
        width = 8 * np.asarray(keys[j]).dtype.itemsize if np.asarray(keys[j]).dtype != np.bool_ else 8
        if descending and descending[j]:
            code = ~code & np.uint64((1 << width) - 1 if width < 64 else 0xFFFFFFFFFFFFFFFF)
        valid = None if valids is None or valids[j] is None else np.asarray(valids[j]).astype(bool)
        if valid is not None:
            code = np.where(valid, code, np.uint64(0))
        perm = perm[np.argsort(code[perm], kind="stable")]
        if valid is not None:
            perm = perm[np.argsort(valid[perm], kind="stable")]      # False (NULL) first
    return perm
