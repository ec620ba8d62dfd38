python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/v5b_pytest.log
python tools/tune.py C4 20000 2000 256x2 288x2 512x1 > gpurun_out/v5b_tune_c4.log 2>&1
python tools/tune.py C3 40000 1000 128x4 256x2 256x3 > gpurun_out/v5b_tune_c3.log 2>&1
GSM_BLOCK=256 GSM_MINB=2 ncu --set full --import-source on --clock-control none -k regex:gsm_lean --launch-skip 3 -c 1 -o gpurun_out/prof_r01_v5b -f python tools/tune.py C4 20000 2000 256x2 > gpurun_out/v5b_ncu.log 2>&1
tail -3 gpurun_out/v5b_pytest.log; cat gpurun_out/v5b_tune_c4.log gpurun_out/v5b_tune_c3.log
