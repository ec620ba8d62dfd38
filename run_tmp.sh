mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
tail -3 gpurun_out/n2_bench.err; cut -c1-700 gpurun_out/n2_bench.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/n2_bench_ref.json 2>> gpurun_out/n2_bench.err
cut -c1-300 gpurun_out/n2_bench_ref.json
