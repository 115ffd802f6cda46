"""Small driver for ncu: one filter + projection pipe over a materialised shard, full scan.
usage: python tools/prof_select.py [rows] [readme|1024|third|all|map] [gen]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fuse_query_b200 import cabi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000_000
case = sys.argv[2] if len(sys.argv) > 2 else "readme"
gen = len(sys.argv) > 3 and sys.argv[3] == "gen"
NUM = "(col number)"
PREDS = {"readme": (f"(< (+ (+ (+ {NUM} (u64 1)) (/ {NUM} (u64 2))) (u64 1)) (u64 100))", 3),
         "1024": (f"(= (* (/ {NUM} (u64 1024)) (u64 1024)) {NUM})", n),
         "third": (f"(= (* (/ {NUM} (u64 3)) (u64 3)) {NUM})", n),
         "all": (f"(>= {NUM} (u64 0))", n),
         "map": (None, n)}
pred, cap = PREDS[case]
ctx = cabi.Context(0)
col = None if gen else ctx.numbers(0, n)
p = ctx.pipe([f"(alias c1 (+ {NUM} (u64 1)))", f"(alias c2 (/ {NUM} (u64 2)))"], predicate=pred, generated=gen)
outs = [ctx.column(cabi.U64, cap), ctx.column(cabi.U64, cap)]
src = cabi.make_source([] if gen else [col], n, generated=gen)
for _ in range(4):
    p.launch_project(src, outs, cap, limit=3 if case == "readme" else -1)
    print(case, p.fetch_project())
