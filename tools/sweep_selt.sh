#!/bin/bash
# Select kernel builds: auto (density probe), forced sparse-tuned, forced dense-tuned
mkdir -p gpurun_out
out=gpurun_out/sweep_selt.log
: > $out
run() { echo "== $*" >> $out; env "$@" TAG="$*" timeout 300 python tools/sweep_select.py >> $out 2>&1; }
run X=auto
run FQ_SEL_VARIANT=sparse
run FQ_SEL_VARIANT=dense
run X=auto
run FQ_SEL_VARIANT=dense FQ_TUNE_SELD_SEG=12
run FQ_SEL_VARIANT=dense FQ_TUNE_SELD_STAGES=12
run FQ_SEL_VARIANT=dense FQ_TUNE_SELD_UNROLL=4 FQ_TUNE_SELD_SEG=4
cat $out
