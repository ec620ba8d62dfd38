mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s22_pytest.log
GSM_LIB_PATH=$PWD/build/libgsm_b200_dbg.so timeout 300 python tools/tune.py C4 20000 2000 256x2 > gpurun_out/s22.log 2>&1
timeout 300 python tools/tune.py C4 20000 2000 256x2 >> gpurun_out/s22.log 2>&1
timeout 300 python tools/tune.py C3 20000 1000 256x2 >> gpurun_out/s22.log 2>&1
timeout 300 python tools/tune.py C2 10000 200 128x4 >> gpurun_out/s22.log 2>&1
cat gpurun_out/s22_pytest.log gpurun_out/s22.log
