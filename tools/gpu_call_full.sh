set -x
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
python tools/prof_select.py 1000000000 > gpurun_out/prof_sel_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fqk_.*select -s 2 -c 1 -o gpurun_out/prof_select_tma_1e9 python tools/prof_select.py 1000000000 > gpurun_out/ncu_sel_tma.log 2>&1
tail -3 gpurun_out/ncu_sel_tma.log | cut -c1-200
