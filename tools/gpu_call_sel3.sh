timeout 900 python -m pytest tests/test_gpu_select_variants.py -m gpu -q -x > gpurun_out/pytest_sel.log 2>&1; tail -3 gpurun_out/pytest_sel.log
TAG=tma python tools/sweep_select.py 2>&1 | tail -1
FQ_SEL_VARIANT=ldg TAG=ldg python tools/sweep_select.py 2>&1 | tail -1
