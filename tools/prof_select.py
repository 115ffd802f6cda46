"""Small driver for ncu: the README filter/projection pipe, full scan, over a materialised shard."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fuse_query_b200 import cabi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000_000
gen = len(sys.argv) > 2 and sys.argv[2] == "gen"
NUM = "(col number)"
ctx = cabi.Context(0)
col = None if gen else ctx.numbers(0, n)
pred = f"(< (+ (+ (+ {NUM} (u64 1)) (/ {NUM} (u64 2))) (u64 1)) (u64 100))"
p = ctx.pipe([f"(alias c1 (+ {NUM} (u64 1)))", f"(alias c2 (/ {NUM} (u64 2)))"], predicate=pred, generated=gen)
outs = [ctx.column(cabi.U64, 3), ctx.column(cabi.U64, 3)]
src = cabi.make_source([] if gen else [col], n, generated=gen)
for _ in range(4):
    p.launch_project(src, outs, 3, limit=3)
    print(p.fetch_project())
