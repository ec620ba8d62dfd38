mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e > gpurun_out/p22_bench_n8.json 2> gpurun_out/p22_bench.err
NCCL_MAX_NCHANNELS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e > gpurun_out/p22_bench_n8_c1.json 2>> gpurun_out/p22_bench.err
cut -c1-260 gpurun_out/p22_bench_n8.json gpurun_out/p22_bench_n8_c1.json; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/p22_bench.err | tail -3
