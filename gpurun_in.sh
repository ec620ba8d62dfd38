mkdir -p gpurun_out
export GSM_LIB_PATH=$PWD/build/libgsm_dev.so
timeout 300 python tools/tune.py C4 20000 2000 128x3 > gpurun_out/p23.log 2>&1
export GSM_LIB_PATH=$PWD/build/libgsm_dev2.so
timeout 300 python tools/tune.py C4 20000 2000 128x3 >> gpurun_out/p23.log 2>&1
unset GSM_LIB_PATH
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> gpurun_out/p23.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/p23_bench_n2.json 2> gpurun_out/p23_bench.err
grep -v "^\[" gpurun_out/p23.log | cut -c1-250; cut -c1-260 gpurun_out/p23_bench_n2.json; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/p23_bench.err | tail -3
