mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s14_pytest.log
GSM_LIB_PATH=$PWD/build/libgsm_b200_dbg.so timeout 600 python tools/tune.py C4 20000 2000 256x2 > gpurun_out/s14.log 2>&1
timeout 600 python tools/tune.py C3 20000 1000 256x2 >> gpurun_out/s14.log 2>&1
timeout 600 python tools/tune.py C2 10000 200 128x4 >> gpurun_out/s14.log 2>&1
GSM_BLOCK=256 GSM_MINB=2 timeout 600 python tools/tune.py C4 20000 2000 256x2 >> gpurun_out/s14.log 2>&1 && GSM_BLOCK=256 GSM_MINB=2 timeout 900 ncu --set full --import-source on --clock-control none -k regex:gsm_lean --launch-skip 3 -c 1 -o gpurun_out/prof_s14_lean -f python tools/tune.py C4 20000 2000 256x2 > gpurun_out/s14_ncu_full.log 2>&1
cat gpurun_out/s14_pytest.log gpurun_out/s14.log
