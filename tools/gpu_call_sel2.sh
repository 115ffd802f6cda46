set -x
export FQ_SEL_VARIANT=tma
python tools/prof_select.py 1000000000 > gpurun_out/prof_sel_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fqk_.*select -s 2 -c 1 -o gpurun_out/prof_select_tma_1e9 python tools/prof_select.py 1000000000 > gpurun_out/ncu_sel_tma.log 2>&1
tail -3 gpurun_out/ncu_sel_tma.log | cut -c1-200
cat > /tmp/b.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from fuse_query_b200 import cabi
n = 1_000_000_000
NUM = "(col number)"
ctx = cabi.Context(0)
stream = torch.cuda.current_stream().cuda_stream
col = ctx.numbers(0, n, stream)
pred = f"(< (+ (+ (+ {NUM} (u64 1)) (/ {NUM} (u64 2))) (u64 1)) (u64 100))"
p = ctx.pipe([f"(alias c1 (+ {NUM} (u64 1)))", f"(alias c2 (/ {NUM} (u64 2)))"], predicate=pred)
outs = [ctx.column(cabi.U64, 3), ctx.column(cabi.U64, 3)]
src = cabi.make_source([col], n)
def run(): p.launch_project(src, outs, 3, limit=3, stream=stream)
for _ in range(3): run()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): run()
b.record(); torch.cuda.synchronize()
print(os.environ.get("TAG"), round(a.elapsed_time(b) / 10, 3), "ms", p.fetch_project())
PY
for cfg in "FQ_TUNE_SELT_THREADS=256" "FQ_TUNE_SELT_THREADS=512" "FQ_TUNE_SELT_THREADS=256 FQ_TUNE_SELT_STAGES=6" "FQ_TUNE_SELT_THREADS=256 FQ_TUNE_SELT_UNROLL=4 FQ_TUNE_SELT_SEG=8" "FQ_TUNE_SELT_THREADS=512 FQ_TUNE_SELT_UNROLL=4 FQ_TUNE_SELT_SEG=8" "FQ_TUNE_SELT_THREADS=384 FQ_TUNE_SELT_UNROLL=4 FQ_TUNE_SELT_SEG=8 FQ_TUNE_SELT_STAGES=8"; do
  env $cfg TAG="$cfg" python /tmp/b.py 2>&1 | tail -1
done
