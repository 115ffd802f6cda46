"""Smallest end-to-end case for compute-sanitizer: every kernel (fill, aggregate, select, map) once."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fuse_query_b200 import cabi
NUM = "(col number)"
ctx = cabi.Context(0)
n = 300_001
col = ctx.numbers(5, n)
p = ctx.pipe([f"(/ (sum {NUM}) (count {NUM}))", f"(max {NUM})", f"(min {NUM})"], aggregate=True)
p.launch_aggregate(cabi.make_source([col], n))
print(p.fetch_aggregate())
pred = f"(= (* (/ {NUM} (u64 3)) (u64 3)) {NUM})"
q = ctx.pipe([NUM, f"(+ {NUM} (u64 1))"], predicate=pred)
outs = [ctx.column(cabi.U64, n), ctx.column(cabi.U64, n)]
q.launch_project(cabi.make_source([col], n), outs, n)
sel, wr = q.fetch_project()
got = outs[0].to_numpy(wr)
exp = np.arange(5, 5 + n, dtype=np.uint64)
exp = exp[exp % 3 == 0]
assert np.array_equal(got, exp), (sel, wr)
q.launch_project(cabi.make_source([col], n), outs, n, limit=7, early_exit=True)
print(q.fetch_project())
m = ctx.pipe([f"(* {NUM} (u64 2))"])
m.launch_project(cabi.make_source([col], n), outs[:1], n)
print(m.fetch_project(), outs[0].to_numpy(3))
print("sanitize case ok")
