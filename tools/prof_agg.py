"""Small driver for ncu: one aggregate pipe over a materialised (or generated) shard.
usage: python tools/prof_agg.py [rows] [headline|sum|max|nullable] [gen]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fuse_query_b200 import cabi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
case = sys.argv[2] if len(sys.argv) > 2 else "headline"
gen = len(sys.argv) > 3 and sys.argv[3] == "gen"
NUM = "(col number)"
EXPRS = {"headline": [f"(/ (sum {NUM}) (count {NUM}))", f"(max {NUM})", f"(min {NUM})"], "sum": [f"(sum {NUM})"], "max": [f"(max {NUM})"],
         "nullable": [f"(sum {NUM})", f"(min {NUM})"]}
ctx = cabi.Context(0)
if case == "nullable":          # UInt64 column with a validity column: traffic = values + validity
    col = ctx.numbers(0, n)
    valid = ctx.from_numpy(np.ones(n, dtype=np.uint8))
    ctx.check(cabi.lib().fq_column_set_validity(ctx._h, col._h, valid._h))
    p = ctx.pipe(EXPRS[case], aggregate=True, nullable=[True])
else:
    col = None if gen else ctx.numbers(0, n)
    p = ctx.pipe(EXPRS[case], aggregate=True, generated=gen)
src = cabi.make_source([] if gen else [col], n, generated=gen)
for _ in range(4):
    p.launch_aggregate(src)
    print(case, p.fetch_aggregate())
