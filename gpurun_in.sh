mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/p19_bench_n2.json 2> gpurun_out/p19_bench.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/p19_bench.json 2>> gpurun_out/p19_bench.err
cut -c1-260 gpurun_out/p19_bench.json gpurun_out/p19_bench_n2.json; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/p19_bench.err | tail -3
