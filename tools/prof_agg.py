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
         }
ctx = cabi.Context(0)
if case.startswith("nullable"):   # nullable[_u8][_bits]: UInt64 / UInt8 values with validity as bytes or as an Arrow bitmap
    import ctypes as C
    u8 = "_u8" in case
    bits = case.endswith("_bits")
    vals = (np.arange(n, dtype=np.uint64) & np.uint64(0xff)).astype(np.uint8) if u8 else None
    col = ctx.from_numpy(vals) if u8 else ctx.numbers(0, n)
    if bits:
        bm = ctx.from_numpy(np.full((n + 7) // 8, 0xff, dtype=np.uint8))
        ctx.check(cabi.lib().fq_column_set_validity_bitmap(ctx._h, col._h, bm._h, 0))
    else:
        flags = np.ones(n, dtype=np.uint8)
        valid = ctx.column(cabi.BOOL, n)
        ctx.check(cabi.lib().fq_column_upload(ctx._h, valid._h, 0, C.c_void_p(flags.ctypes.data), n, None))
        ctx.synchronize()
        ctx.check(cabi.lib().fq_column_set_validity(ctx._h, col._h, valid._h))
    p = ctx.pipe(["(sum (col number))", "(min (col number))"], aggregate=True, dtypes=[cabi.U8 if u8 else cabi.U64], nullable=[2 if bits else 1])
else:
    col = None if gen else ctx.numbers(0, n)
    p = ctx.pipe(EXPRS[case], aggregate=True, generated=gen)
src = cabi.make_source([] if gen else [col], n, generated=gen)
for _ in range(4):
    p.launch_aggregate(src)
    print(case, p.fetch_aggregate())
