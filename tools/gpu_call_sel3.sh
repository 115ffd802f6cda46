timeout 900 python -m pytest tests/test_gpu_select_variants.py tests/test_host_gpu.py -m gpu -q -x -k "without_filter or mysql or parquet" > gpurun_out/pytest_sel.log 2>&1; tail -5 gpurun_out/pytest_sel.log
TAG=map-ldg python tools/sweep_select.py 2>&1 | tail -1
FQ_MAP_VARIANT=tma TAG=map-tma python tools/sweep_select.py 2>&1 | tail -1
FQ_MAP_VARIANT=tma FQ_TUNE_TMA_STAGES=3 TAG=map-tma-3 python tools/sweep_select.py 2>&1 | tail -1
FQ_MAP_VARIANT=tma FQ_TUNE_TMA_STAGES=6 TAG=map-tma-6 python tools/sweep_select.py 2>&1 | tail -1
FQ_MAP_VARIANT=tma FQ_TUNE_TMA_THREADS=512 FQ_TUNE_TMA_UNROLL=4 TAG=map-tma-512x4 python tools/sweep_select.py 2>&1 | tail -1
