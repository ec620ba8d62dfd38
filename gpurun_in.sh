mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/p8_pytest.log
cat gpurun_out/p8_pytest.log
