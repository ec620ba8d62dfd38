mkdir -p gpurun_out
timeout 600 python tools/sanitize_case.py > gpurun_out/p28.log 2>&1; tail -3 gpurun_out/p28.log
