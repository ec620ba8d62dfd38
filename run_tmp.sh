mkdir -p gpurun_out
timeout 90 python tools/tune.py C4 60000 2000 128x3n 128x3 128x3n 128x3 > gpurun_out/p1_tune_c4.log 2>&1 || { echo HANG_OR_FAIL; tail -3 gpurun_out/p1_tune_c4.log; exit 1; }
timeout 90 python tools/tune.py C3 50000 1000 128x3n 128x3 > gpurun_out/p1_tune_c3.log 2>&1
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/p1_pytest.log
cat gpurun_out/p1_tune_c4.log gpurun_out/p1_tune_c3.log gpurun_out/p1_pytest.log
