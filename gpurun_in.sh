set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s2_pytest.log
timeout 600 python tools/tune.py C4 20000 2000 256x2 > gpurun_out/s2_tune_c4.log 2>&1
timeout 600 python tools/tune.py C3 20000 1000 256x2 256x3 > gpurun_out/s2_tune_c3.log 2>&1
timeout 600 python tools/tune.py C2 10000 200 128x6 128x4 > gpurun_out/s2_tune_c2.log 2>&1
tail -3 gpurun_out/s2_pytest.log; cat gpurun_out/s2_tune_c4.log gpurun_out/s2_tune_c3.log gpurun_out/s2_tune_c2.log
