timeout 900 python bench.py --workload c4full --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/r2f_c4full.json 2> gpurun_out/r2f_c4full.err
tail -3 gpurun_out/r2f_c4full.err; cut -c1-900 gpurun_out/r2f_c4full.json
nvidia-smi --query-gpu=memory.total,memory.used --format=csv
