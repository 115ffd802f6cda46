# usage: bash tools/gpu_call_merge.sh N   -- headline step with the in-kernel peer merge and with NCCL
N=$1
port=29600
for m in peer nccl; do
  port=$((port + 1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 50 --warmup 3 --merge $m --no-query-table --e2e-rows 160000000 --e2e-steps 1 > gpurun_out/bench_n${N}_$m.json 2> gpurun_out/bench_n${N}_$m.err
  grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/bench_n${N}_$m.err | tail -3
  python - <<PY
import json
lines = [l for l in open("gpurun_out/bench_n${N}_$m.json") if l.startswith("{")]
if lines:
    d = json.loads(lines[-1])
    print("$m", d["ms_per_step"], d["roofline"]["kernel_ms"], d["value"], d["config"]["merge"][:70], d["result"])
else:
    print("$m: no JSON line")
PY
done
