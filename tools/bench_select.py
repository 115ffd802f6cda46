"""Micro-benchmark of the select (filter + projection + limit) and map kernels — tuning aid.
usage: python tools/bench_select.py [rows]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fuse_query_b200 import cabi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
NUM = "(col number)"
ctx = cabi.Context(0)
stream = torch.cuda.current_stream().cuda_stream
col = ctx.numbers(0, n, stream)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


cases = {
    "readme (66 of n)": (f"(< (+ (+ (+ {NUM} (u64 1)) (/ {NUM} (u64 2))) (u64 1)) (u64 100))", 3),
    "1/1024 selected": (f"(= (* (/ {NUM} (u64 1024)) (u64 1024)) {NUM})", -1),
    "1/3 selected": (f"(= (* (/ {NUM} (u64 3)) (u64 3)) {NUM})", -1),
    "all selected": (f"(>= {NUM} (u64 0))", -1),
}
proj = [f"(alias c1 (+ {NUM} (u64 1)))", f"(alias c2 (/ {NUM} (u64 2)))"]
res = {}
for name, (pred, limit) in cases.items():
    for gen in (False, True):
        p = ctx.pipe(proj, predicate=pred, generated=gen)
        cap = 3 if limit == 3 else n
        outs = [ctx.column(cabi.U64, cap), ctx.column(cabi.U64, cap)]
        src = cabi.make_source([] if gen else [col], n, generated=gen)
        for early in ((False, True) if limit >= 0 else (False,)):
            ms = timeit(lambda: p.launch_project(src, outs, cap, limit=limit, early_exit=early, stream=stream))
            sel, wr = p.fetch_project()
            byts = (0 if gen else 8) * n + 16 * wr
            res[f"{name} | {'gen' if gen else 'mat'} | early={early}"] = (round(ms, 3), sel, round(byts / ms / 1e6), "GB/s")
        for o in outs:
            o.free()
        p.destroy()
p = ctx.pipe(proj)
outs = [ctx.column(cabi.U64, n), ctx.column(cabi.U64, n)]
ms = timeit(lambda: p.launch_project(cabi.make_source([col], n), outs, n, stream=stream))
res["map 2 exprs | mat"] = (round(ms, 3), n, round(24 * n / ms / 1e6), "GB/s")
for k, v in res.items():
    print(f"{k:50s} {v}")
