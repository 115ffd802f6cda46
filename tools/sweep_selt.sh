#!/bin/bash
# Shape / cache-policy sweep of the select kernels (NVRTC builds with other #defines)
mkdir -p gpurun_out
out=gpurun_out/sweep_selt.log
: > $out
run() { echo "== $*" >> $out; env "$@" TAG="$*" timeout 300 python tools/sweep_select.py >> $out 2>&1; }
run X=default
run "FQ_TUNE_EXTRA=#define FQ_STORE_CS 0"
run "FQ_TUNE_EXTRA=#define FQ_L2_HINTS 0"
run "FQ_TUNE_EXTRA=#define FQ_STORE_CS 0;#define FQ_L2_HINTS 0"
run FQ_TUNE_SELT_SEG=4
run FQ_TUNE_SELT_SEG=4 FQ_TUNE_SELT_LAG=2
run FQ_TUNE_SELT_SEG=8 FQ_TUNE_SELT_LAG=2
run FQ_TUNE_SELT_SEG=4 "FQ_TUNE_EXTRA=#define FQ_STORE_CS 0"
run FQ_TUNE_SELT_SEG=2 FQ_TUNE_SELT_LAG=3
run FQ_SELT_STAGE2=0
run GEN=1
run GEN=1 FQ_TUNE_SEL_THREADS=352
run GEN=1 FQ_TUNE_SEL_THREADS=352 FQ_TUNE_SEL_MIN_BLOCKS=2
run GEN=1 FQ_TUNE_SEL_THREADS=480 FQ_TUNE_SEL_MIN_BLOCKS=1
run GEN=1 FQ_TUNE_SEL_THREADS=224 FQ_TUNE_SEL_MIN_BLOCKS=3
cat $out
