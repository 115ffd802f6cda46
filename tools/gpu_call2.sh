set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
SMALL="--rows 2000000000 --steps 3 --warmup 3 --no-cpu-baseline --no-query-table --e2e-rows 160000000 --e2e-steps 1"
for U in 4 8; do for B in 0 2 3; do
  FQ_AGG_UNROLL=$U FQ_AGG_BLOCKS_PER_SM=$B python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-query-table --e2e-rows 160000000 --e2e-steps 1 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TUNE U=$U B=$B', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done; done
python bench.py $SMALL > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py $SMALL > gpurun_out/ncu_list.log 2>&1
python bench.py $SMALL > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fqk_ -s 3 -c 2 -o gpurun_out/prof_agg python bench.py $SMALL > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
