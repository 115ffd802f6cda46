#!/bin/bash
# duration + DRAM bytes + instruction count of one launch per case (metrics-only ncu: cheap)
mkdir -p gpurun_out
out=gpurun_out/quick_prof.log
: > $out
m() { name=$1; regex=$2; shift 2; echo "== $name $FQ_GB_NOAGG" >> $out; timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:$regex --launch-skip 2 --launch-count 1 --csv "$@" 2>/dev/null | grep -E "dram__bytes|gpu__time|inst_executed" | awk -F'","' '{print "   ", $(NF-2), $(NF-1), $NF}' >> $out; }
for k in 7 1000 100000 1000000 100000000; do m groupby_k$k groupby python tools/prof_groupby.py 1000000000 $k; done
export FQ_TUNE_EXTRA="#define FQ_GB_NO_WARP_AGG 1"; export FQ_GB_NOAGG=noagg
for k in 7 1000 1000000; do m groupby_k$k groupby python tools/prof_groupby.py 1000000000 $k; done
cat $out
