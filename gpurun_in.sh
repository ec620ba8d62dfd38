mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/f1_bench.json 2> gpurun_out/f1_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f1_bench_ref.json 2>> gpurun_out/f1_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f1_smoke.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gsm_ -c 40 --csv --log-file gpurun_out/f1_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/f1_ncu_list.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gsm_lean --launch-skip 3 -c 1 -o gpurun_out/prof_f1_lean -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/f1_ncu_full.log 2>&1
cut -c1-300 gpurun_out/f1_bench.json; cut -c1-200 gpurun_out/f1_bench_ref.json; cat gpurun_out/f1_smoke.log | tail -2; grep real gpurun_out/f1_bench.err
