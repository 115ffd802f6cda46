# final ncu evidence for the headline command (run only after the same command exited 0 without ncu)
ARGS="--steps 2 --warmup 3 --no-cpu-baseline --no-query-table --e2e-rows 160000000 --e2e-steps 1"
python bench.py $ARGS > gpurun_out/plain_full.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fqk_.*agg -s 3 -c 1 -o gpurun_out/prof_agg_final_1e10 python bench.py $ARGS > gpurun_out/ncu_full_1e10.log 2>&1
tail -2 gpurun_out/ncu_full_1e10.log | cut -c1-200
python bench.py $ARGS > gpurun_out/plain_full2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_final_1e10.csv python bench.py $ARGS > gpurun_out/ncu_list_1e10.log 2>&1
tail -2 gpurun_out/ncu_list_1e10.log | cut -c1-200
