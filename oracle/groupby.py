"""GROUP BY oracle — TEST INFRASTRUCTURE ONLY (tests/, never the product path).

The reference plans GROUP BY (planners/plan_parser.rs:279-308: AggregatePlan{group_expr, aggr_expr}, schema = group fields then
aggregate fields, planners/plan_builder.rs:63-83) but its pipeline builder only uses aggr_expr (processors/pipeline_builder.rs:
50-65): the operator is never executed there, so there is no reference behaviour to pin — PARITY UNPINNED for this operator.
What this file states is the operator that plan describes, built from pieces that ARE pinned: the rows are split by the
tuple of key values (evaluated by the C oracle's Function::eval; a NULL key is a group of its own, as in SQL) and every
group runs the reference's own aggregate protocol — accumulate -> accumulate_result -> merge_state -> merge_result
(functions/function_aggregator.rs:57-143, function_arithmetic.rs:69-88) — through the C oracle, exactly as an un-grouped
query over just those rows would.  tests/test_oracle_groupby.py cross-checks it against pyarrow's group_by.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import binding as o


def _take(a: "o.Array", idx: np.ndarray) -> "o.Array":
    return o.Array(a.dtype, np.ascontiguousarray(a.values[idx]), None if a.valid is None else np.ascontiguousarray(a.valid[idx]))


def run_group_by(keys: Sequence[str], aggs: Sequence[str], *, total: Optional[int] = None, table: Optional[dict] = None,
                 predicate: Optional[str] = None) -> Tuple[List[str], List[tuple]]:
    """-> (column names: keys then aggregates, rows sorted by key tuple with NULL first).  s-expressions as in fq_oracle.h."""
    cols = dict(table) if table is not None else {"number": o.from_numpy(np.arange(total, dtype=np.uint64))}
    n = len(next(iter(cols.values())))
    idx = np.arange(n)
    if predicate is not None:
        m = o.Function(predicate).eval(cols)
        keep = m.values.astype(bool)
        if m.valid is not None:
            keep &= m.valid.astype(bool)          # a NULL predicate slot keeps nothing
        idx = idx[keep]
    sub = {k: _take(a, idx) for k, a in cols.items()}
    key_fns = [o.Function(k) for k in keys]
    key_arrays = [f.eval(sub) for f in key_fns]
    names = [f.display() for f in key_fns] + [o.Function(a).display() for a in aggs]
    if len(idx) == 0:
        return names, []
    # sort rows by (null flag, value) of every key, last key fastest
    sort_cols = []
    for a in reversed(key_arrays):
        v = a.values
        if a.valid is not None:
            v = np.where(a.valid.astype(bool), v, np.zeros(1, dtype=v.dtype))
            sort_cols += [v, a.valid.astype(np.uint8)]
        else:
            sort_cols.append(v)
    order = np.lexsort(sort_cols)
    change = np.zeros(len(order), dtype=bool)
    change[0] = True
    for a in key_arrays:
        v = a.values[order]
        ok = np.ones(len(order), dtype=bool) if a.valid is None else a.valid[order].astype(bool)
        v = np.where(ok, v, np.zeros(1, dtype=v.dtype))
        if v.dtype.kind == "f":
            v = v.view(np.uint32 if v.dtype.itemsize == 4 else np.uint64)      # groups are by bit pattern
        change[1:] |= (v[1:] != v[:-1]) | (ok[1:] != ok[:-1])
    starts = np.flatnonzero(change)
    ends = np.append(starts[1:], len(order))
    rows = []
    for s, e in zip(starts, ends):
        g = order[s:e]
        g.sort()                                   # original row order inside the group (float sums follow it)
        block = {k: _take(a, g) for k, a in sub.items()}
        key = tuple(a.to_list()[g[0]] for a in key_arrays)
        vals = []
        for a in aggs:
            partial = o.Function(a)
            partial.accumulate(block)
            final = o.Function(a)
            final.merge_state(partial.accumulate_result())
            v = final.merge_result()
            vals.append(v.value)
        rows.append(key + tuple(vals))
    rows.sort(key=lambda r: tuple((x is not None, x) for x in r[:len(keys)]))
    return names, rows
