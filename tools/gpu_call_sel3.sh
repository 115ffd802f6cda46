timeout 900 python -m pytest tests/test_gpu_select_variants.py -m gpu -q -x > gpurun_out/pytest_sel.log 2>&1; tail -3 gpurun_out/pytest_sel.log
export FQ_SEL_VARIANT=tma
for cfg in "X=1" "FQ_TUNE_SELT_LAG=2" "FQ_TUNE_SELT_LAG=1" "FQ_TUNE_SELT_LAG=3 FQ_TUNE_SELT_STAGES=4" "FQ_TUNE_SELT_LAG=2 FQ_TUNE_SELT_STAGES=4" "FQ_TUNE_SELT_LAG=4"; do
  env $cfg TAG="$cfg" python tools/sweep_select.py 2>&1 | tail -1
done
