python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2e_n8_p2p.json 2> gpurun_out/r2e_n8_p2p.err
tail -2 gpurun_out/r2e_n8_p2p.err
python -c "
import json
d=json.load(open('gpurun_out/r2e_n8_p2p.json'))
print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['gather'][:40], d['config']['gather_check'], d['e2e']['value'])
"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 --gather nccl --no-e2e > gpurun_out/r2e_n8_nccl.json 2> gpurun_out/r2e_n8_nccl.err
python -c "
import json
d=json.load(open('gpurun_out/r2e_n8_nccl.json'))
print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['gather'][:40])
"
python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2e_n1.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r2e_n1.json'))
print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'])
"
nvidia-smi topo -m 2>/dev/null | head -14; lscpu | grep -i "numa\|socket\|model name" | head
