"""Shape sweep aid for the select kernels: README / sparse / dense selections over 1e9 materialised rows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fuse_query_b200 import cabi
n = 1_000_000_000
NUM = "(col number)"
ctx = cabi.Context(0)
stream = torch.cuda.current_stream().cuda_stream
gen = os.environ.get("GEN") == "1"      # generated source: the LDG select kernel
col = ctx.numbers(0, n, stream)
src = cabi.make_source([] if gen else [col], n, generated=gen)
proj = [f"(alias c1 (+ {NUM} (u64 1)))", f"(alias c2 (/ {NUM} (u64 2)))"]
cases = {"readme": (f"(< (+ (+ (+ {NUM} (u64 1)) (/ {NUM} (u64 2))) (u64 1)) (u64 100))", 3),
         "1/1024": (f"(= (* (/ {NUM} (u64 1024)) (u64 1024)) {NUM})", n),
         "1/3": (f"(= (* (/ {NUM} (u64 3)) (u64 3)) {NUM})", n),
         "all": (f"(>= {NUM} (u64 0))", n)}
outs = [ctx.column(cabi.U64, n), ctx.column(cabi.U64, n)]
res = []
for name, (pred, cap) in cases.items():
    p = ctx.pipe(proj, predicate=pred, generated=gen)
    run = lambda: p.launch_project(src, outs, cap, stream=stream)
    for _ in range(2): run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): run()
    b.record(); torch.cuda.synchronize()
    res.append(f"{name} {a.elapsed_time(b) / 5:.3f}")
p = ctx.pipe(proj, generated=gen)
run = lambda: p.launch_project(src, outs, n, stream=stream)
for _ in range(2): run()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): run()
b.record(); torch.cuda.synchronize()
res.append(f"map {a.elapsed_time(b) / 5:.3f}")
print(os.environ.get("TAG", ""), "|", " | ".join(res))
