mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/p13_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/p13_bench.json 2> gpurun_out/p13_bench.err
timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/p13_bench_c3.json 2>> gpurun_out/p13_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gsm_ -c 40 --csv --log-file gpurun_out/p13_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/p13_ncu_list.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gsm_lean --launch-skip 3 -c 1 -o gpurun_out/prof_p13_lean -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/p13_ncu_full.log 2>&1
cat gpurun_out/p13_pytest.log; cut -c1-260 gpurun_out/p13_bench.json gpurun_out/p13_bench_c3.json; tail -3 gpurun_out/p13_bench.err
