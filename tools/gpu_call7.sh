timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "compaction or filter or projection or capacity or mixed" 2>&1 | tail -2
echo "== default"; timeout 300 python tools/bench_select.py 2>&1 | tail -11
for cfg in "FQ_TUNE_SEL_THREADS=256 FQ_TUNE_SEL_MIN_BLOCKS=3" "FQ_TUNE_SEL_THREADS=256 FQ_TUNE_SEL_MIN_BLOCKS=2" "FQ_TUNE_SEL_THREADS=512 FQ_TUNE_SEL_MIN_BLOCKS=1" "FQ_TUNE_SEL_THREADS=384 FQ_TUNE_SEL_MIN_BLOCKS=1"; do
  echo "== $cfg"; env $cfg timeout 300 python tools/bench_select.py 2>&1 | grep -E "early=False|map" | grep -E "mat|gen . early=False" | grep -v "1/1024"
done
