"""N > 1 host logic on CPU: world_size-2 gloo run of the shard -> evaluate -> gather path.
The per-shard evaluation is stood in for by the oracle here (this is a test of the sharding/gather plumbing;
the GPU path itself is covered by the -m gpu tier and by bench.py under torchrun)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gammaspikemodel_b200 import synthetic
from gammaspikemodel_b200.sharding import gather_rows, shard_range
from helpers import host_arrays, run_oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_trees, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = synthetic.make_batch("C3", n_trees, 12, seed=77)
    a = host_arrays(b)
    lo, hi = shard_range(n_trees, rank, world)
    shard = {k: v[lo:hi] for k, v in a.items()}
    local = run_oracle(shard, b.flags, b.stub_mode, rates=False, threads=1)["out"] if hi > lo else np.zeros((0, 8))
    full = gather_rows(torch.from_numpy(local), n_trees_total=n_trees)
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_every_tree_once():
    for T in (0, 1, 7, 8, 9, 1000):
        for G in (1, 2, 4, 8):
            r = [shard_range(T, k, G) for k in range(G)]
            assert r[0][0] == 0 and r[-1][1] == T
            assert all(r[k][1] == r[k + 1][0] for k in range(G - 1))


def _run(n_trees):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_trees, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    b = synthetic.make_batch("C3", n_trees, 12, seed=77)
    ref = run_oracle(host_arrays(b), b.flags, b.stub_mode, rates=False, threads=1)["out"]
    assert np.array_equal(got, ref, equal_nan=True)


def test_two_rank_gather_equal_shards():
    _run(10)


def test_two_rank_gather_ragged_shards():
    _run(7)
