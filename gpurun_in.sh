# standard validation sequence on a GPU box: gpurun --timeout 2400 -- 'bash gpurun_in.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/std_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/std_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/std_bench.json 2> gpurun_out/std_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/std_bench_ref.json 2>> gpurun_out/std_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gsm_ -c 40 --csv --log-file gpurun_out/std_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/std_ncu_list.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gsm_lean --launch-skip 3 -c 1 -o gpurun_out/prof_std_lean -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/std_ncu_full.log 2>&1
cat gpurun_out/std_pytest.log gpurun_out/std_smoke.log; cut -c1-300 gpurun_out/std_bench.json gpurun_out/std_bench_ref.json; tail -3 gpurun_out/std_bench.err
