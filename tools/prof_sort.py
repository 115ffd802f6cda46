"""Small driver for ncu: ORDER BY over n rows.  usage: python tools/prof_sort.py [rows] [seq|scrambled]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fuse_query_b200 import cabi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else "scrambled"
NUM = "(col number)"
ctx = cabi.Context(0)
col = ctx.numbers(0, n)
key = col
if kind == "scrambled":
    p = ctx.pipe([f"(* {NUM} (u64 {0x9E3779B97F4A7C15}))"])
    key = ctx.column(cabi.U64, n)
    p.launch_project(cabi.make_source([col], n), [key], n)
    assert p.fetch_project()[1] == n
for _ in range(2):
    t0 = time.time()
    idx = ctx.sort_indices([key], n, [kind != "scrambled"])
    ctx.synchronize()
    print(kind, n, "rows sorted in", round((time.time() - t0) * 1e3, 2), "ms")
    idx.free()
