mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r1b_bench_n2.json 2> gpurun_out/r1b_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r1b_bench_ref_n2.json 2>> gpurun_out/r1b_bench_n2.err
cat gpurun_out/r1b_bench_n2.json gpurun_out/r1b_bench_ref_n2.json; tail -5 gpurun_out/r1b_bench_n2.err
