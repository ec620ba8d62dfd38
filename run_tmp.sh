mkdir -p gpurun_out
: > gpurun_out/p10.log
GSM_LIB_PATH=$PWD/build/libA.so timeout 90 python tools/tune.py C4 60000 2000 128x3 128x3n pk 128x3 >> gpurun_out/p10.log 2>&1
GSM_LIB_PATH=$PWD/build/libA.so timeout 90 python tools/tune.py C3 50000 1000 128x3 128x3n pk >> gpurun_out/p10.log 2>&1
cut -c1-175 gpurun_out/p10.log
