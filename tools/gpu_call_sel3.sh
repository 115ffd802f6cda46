timeout 900 python -m pytest tests/test_gpu_select_variants.py -m gpu -q -x > gpurun_out/pytest_sel.log 2>&1; tail -3 gpurun_out/pytest_sel.log
export FQ_SEL_VARIANT=tma
for cfg in "X=1" "FQ_TUNE_SELT_STAGES=5" "FQ_TUNE_SELT_THREADS=640" "FQ_TUNE_SELT_THREADS=448 FQ_TUNE_SELT_STAGES=6"; do
  env $cfg TAG="$cfg" python tools/sweep_select.py 2>&1 | tail -1
done
FQ_SEL_VARIANT=ldg TAG=ldg python tools/sweep_select.py 2>&1 | tail -1
FQ_SEL_VARIANT=ldg FQ_TUNE_SEL_LOOK=10 TAG=ldg-look10 python tools/sweep_select.py 2>&1 | tail -1
