"""Wide-row tables: aggregate / select / projection throughput over a 4-column table (u64, i64, f64, i32 = 28 B/row)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fuse_query_b200 import cabi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
ctx = cabi.Context(0)
stream = torch.cuda.current_stream().cuda_stream
names, dts = ["a", "b", "e", "c"], [cabi.U64, cabi.I64, cabi.F64, cabi.I32]
cols = [ctx.numbers(0, n, stream)]
e = ctx.column(cabi.F64, n)
p = ctx.pipe(["(* (col a) (f64 0.5))"], columns=["a"], dtypes=[cabi.U64])   # e = a / 2 as Float64, on the device
p.launch_project(cabi.make_source(cols, n), [e], n, stream=stream)
p.fetch_project()
import ctypes as C
L = cabi.lib()
step = 1 << 24
bcol = ctx.column(cabi.I64, n); ccol = ctx.column(cabi.I32, n)
for off in range(0, n, step):
    m = min(step, n - off)
    hb = np.arange(off, off + m, dtype=np.int64) - 1000
    hc = (np.arange(off, off + m, dtype=np.int64) % 30000).astype(np.int32)
    ctx.check(L.fq_column_upload(ctx._h, bcol._h, off, C.c_void_p(hb.ctypes.data), m, C.c_void_p(stream))); ctx.synchronize(stream)
    ctx.check(L.fq_column_upload(ctx._h, ccol._h, off, C.c_void_p(hc.ctypes.data), m, C.c_void_p(stream))); ctx.synchronize(stream)
cols = [cols[0], bcol, e, ccol]
src = cabi.make_source(cols, n)

def timeit(fn, reps=5):
    for _ in range(2): fn()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b_.record(); torch.cuda.synchronize()
    return a.elapsed_time(b_) / reps

kw = dict(columns=names, dtypes=dts)
agg = ctx.pipe(["(sum (col a))", "(min (col b))", "(max (col e))", "(sum (col c))"], aggregate=True, **kw)
ms = timeit(lambda: agg.launch_aggregate(src, stream=stream))
print(f"aggregate 4 cols (28 B/row): {ms:.3f} ms  {28 * n / ms / 1e6:.0f} GB/s  variant={os.environ.get('FQ_AGG_VARIANT', 'default')}")
agg2 = ctx.pipe(["(sum (col a))", "(max (col c))"], aggregate=True, **kw)
ms = timeit(lambda: agg2.launch_aggregate(src, stream=stream))
print(f"aggregate 2 cols (12 B/row): {ms:.3f} ms  {12 * n / ms / 1e6:.0f} GB/s")
sel = ctx.pipe(["(col a)", "(col e)"], predicate="(and (< (col c) (i32 3)) (> (col b) (i64 0)))", **kw)
outs = [ctx.column(cabi.U64, n // 100), ctx.column(cabi.F64, n // 100)]
ms = timeit(lambda: sel.launch_project(src, outs, n // 100, stream=stream))
print(f"select 1e-4 of rows, pred over 2 cols (12 B/row read): {ms:.3f} ms  {12 * n / ms / 1e6:.0f} GB/s", sel.fetch_project())
