mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/p7_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/p7_bench.json 2> gpurun_out/p7_bench.err
timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/p7_bench_c3.json 2>> gpurun_out/p7_bench.err
timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/p7_bench_c2.json 2>> gpurun_out/p7_bench.err
cat gpurun_out/p7_pytest.log; cut -c1-600 gpurun_out/p7_bench*.json; tail -3 gpurun_out/p7_bench.err
