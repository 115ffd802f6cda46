#!/bin/bash
# Shape sweep of the staged select kernel with DRAM traffic: per shape the timings of the four selections, then
# ncu's dram bytes + duration of one dense ("all") launch
mkdir -p gpurun_out
out=gpurun_out/sweep_selt.log
: > $out
run() {
  echo "== $*" >> $out
  env "$@" TAG="$*" timeout 300 python tools/sweep_select.py >> $out 2>&1
  env "$@" timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_op_read_hit_rate.pct --clock-control none -k regex:select --launch-skip 2 --launch-count 1 --csv python tools/prof_select.py 1000000000 all 2>/dev/null | grep -E "dram__bytes|gpu__time|hit_rate" | awk -F'","' '{print "   ncu", $(NF-2), $(NF-1), $NF}' >> $out
}
run X=default
run FQ_TUNE_SELT_LAG=1
run FQ_TUNE_SELT_SEG=6 FQ_TUNE_SELT_LAG=2
run FQ_TUNE_SELT_UNROLL=2 FQ_TUNE_SELT_SEG=12
run FQ_TUNE_SELT_UNROLL=2 FQ_TUNE_SELT_SEG=16
run FQ_TUNE_SELT_UNROLL=2 FQ_TUNE_SELT_SEG=16 FQ_TUNE_SELT_LAG=2
run FQ_TUNE_SELT_UNROLL=2 FQ_TUNE_SELT_SEG=8
cat $out
