mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/p11_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/p11_bench.json 2> gpurun_out/p11_bench.err
cat gpurun_out/p11_pytest.log; cut -c1-250 gpurun_out/p11_bench.json; tail -3 gpurun_out/p11_bench.err
