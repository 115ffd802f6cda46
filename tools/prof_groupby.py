"""Small driver for ncu: GROUP BY number % k with sum / count / min / max over a materialised shard.
usage: python tools/prof_groupby.py [rows] [k]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fuse_query_b200 import cabi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
NUM = "(col number)"
ctx = cabi.Context(0)
col = ctx.numbers(0, n)
p = ctx.pipe([f"(sum {NUM})", f"(count {NUM})", f"(min {NUM})", f"(max {NUM})"], keys=[f"(- {NUM} (* (/ {NUM} (u64 {k})) (u64 {k})))"])
p.groupby_reserve(min(k, n))
src = cabi.make_source([col], n)
for _ in range(3):
    p.launch_groupby(src)
    print(k, p.fetch_groupby())
