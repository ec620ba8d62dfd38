mkdir -p gpurun_out
export GSM_LIB_PATH=$PWD/build/libgsm_dev.so
timeout 300 python tools/tune.py C3 20000 1000 128x3 > gpurun_out/p27.log 2>&1
export GSM_LIB_PATH=$PWD/build/libgsm_dev2.so
timeout 300 python tools/tune.py C3 20000 1000 128x3 >> gpurun_out/p27.log 2>&1
unset GSM_LIB_PATH
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> gpurun_out/p27.log
timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/p27_bench_c3.json 2> gpurun_out/p27.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/p27_bench.json 2>> gpurun_out/p27.err
grep -v "^\[" gpurun_out/p27.log | cut -c1-200; cut -c1-200 gpurun_out/p27_bench_c3.json gpurun_out/p27_bench.json; tail -2 gpurun_out/p27.err
