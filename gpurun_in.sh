mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e > gpurun_out/p24_bench_n8.json 2> gpurun_out/p24_bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29592 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/p24_bench_n8_full.json 2>> gpurun_out/p24_bench.err
cut -c1-260 gpurun_out/p24_bench_n8.json gpurun_out/p24_bench_n8_full.json; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/p24_bench.err | tail -3
