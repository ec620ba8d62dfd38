mkdir -p gpurun_out
: > gpurun_out/p12.log
timeout 90 python tools/tune.py C4 60000 2000 128x3 >> gpurun_out/p12.log 2>&1
GSM_ROOT_SWAP=1 timeout 90 python tools/tune.py C4 60000 2000 128x3 >> gpurun_out/p12.log 2>&1
cut -c1-150 gpurun_out/p12.log
