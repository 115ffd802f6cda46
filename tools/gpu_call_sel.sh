set -x
timeout 900 python -m pytest tests/test_gpu_select_variants.py -m gpu -q -x > gpurun_out/pytest_sel.log 2>&1; tail -15 gpurun_out/pytest_sel.log
for v in ldg tma; do
  FQ_SEL_VARIANT=$v timeout 300 python tools/bench_select.py 1000000000 > gpurun_out/bench_select_$v.log 2>&1; cat gpurun_out/bench_select_$v.log
done
