#!/bin/bash
# ncu one-liners for the GROUP BY kernel: KS="7 1000" TAGX=label bash tools/quick_prof.sh  (appends to gpurun_out/quick_prof.log)
mkdir -p gpurun_out
out=gpurun_out/quick_prof.log
M=dram__bytes_read.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__inst_executed_op_global_red.sum,smsp__inst_executed_op_shared_atom.sum
m() { name=$1; regex=$2; shift 2; echo "== $name $TAGX" >> $out; timeout 300 ncu --metrics $M --clock-control none -k regex:$regex --launch-skip 2 --launch-count 1 --csv "$@" 2>/dev/null | grep -E '^"[0-9]' | awk -F'","' '{print "   ", $(NF-2), $(NF-1), $NF}' >> $out; }
for k in ${KS:-7 100 1000 1500 5000 1000000}; do m groupby_k$k groupby python tools/prof_groupby.py 1000000000 $k; done
cat $out
