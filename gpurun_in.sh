set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/s1_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s1_pytest.log
timeout 600 python tools/tune.py C4 20000 2000 256x2 288x2 384x2 512x1 > gpurun_out/s1_tune_c4.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/s1_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gsm_ -c 60 --csv --log-file gpurun_out/s1_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s1_ncu_list.log 2>&1
GSM_BLOCK=256 GSM_MINB=2 timeout 900 ncu --set full --import-source on --clock-control none -k regex:gsm_lean --launch-skip 3 -c 1 -o gpurun_out/prof_s1_lean -f python tools/tune.py C4 20000 2000 256x2 > gpurun_out/s1_ncu_full.log 2>&1
tail -3 gpurun_out/s1_pytest.log; cat gpurun_out/s1_tune_c4.log; tail -2 gpurun_out/s1_bench.log
