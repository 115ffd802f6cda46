#!/bin/bash
# duration + DRAM bytes of one launch per case (metrics-only ncu: cheap)
mkdir -p gpurun_out
out=gpurun_out/quick_prof.log
: > $out
m() { name=$1; regex=$2; shift 2; echo "== $name" >> $out; timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$regex --launch-skip 2 --launch-count 1 --csv "$@" 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' '{print "   ", $(NF-2), $(NF-1), $NF}' >> $out; }
m nullable_u64_bytes "_agg_" python tools/prof_agg.py 1000000000 nullable
m nullable_u64_bits "_agg_" python tools/prof_agg.py 1000000000 nullable_bits
m nullable_u8_bytes "_agg_" python tools/prof_agg.py 4000000000 nullable_u8
m nullable_u8_bits "_agg_" python tools/prof_agg.py 4000000000 nullable_u8_bits
m groupby_k7 groupby python tools/prof_groupby.py 1000000000 7
m groupby_k1000 groupby python tools/prof_groupby.py 1000000000 1000
m groupby_k1e6 groupby python tools/prof_groupby.py 1000000000 1000000
cat $out
