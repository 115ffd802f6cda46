#!/bin/bash
# Shape / cache-policy sweep of the select kernels (NVRTC builds with other #defines)
mkdir -p gpurun_out
out=gpurun_out/sweep_selt.log
: > $out
run() { echo "== $*" >> $out; env "$@" TAG="$*" timeout 300 python tools/sweep_select.py >> $out 2>&1; }
run X=default
run FQ_TUNE_SELT_STAGES=7
run FQ_TUNE_SELT_SEG=6
run FQ_TUNE_SELT_SEG=6 FQ_TUNE_SELT_STAGES=7
run FQ_TUNE_SELT_SEG=4
run FQ_TUNE_SELT_SEG=6 FQ_TUNE_SELT_LAG=2
run FQ_TUNE_SELT_SEG=4 FQ_TUNE_SELT_LAG=4
run FQ_TUNE_SELT_UNROLL=2 FQ_TUNE_SELT_SEG=12
cat $out
