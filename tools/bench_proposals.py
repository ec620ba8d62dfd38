#!/usr/bin/env python
"""BASELINE config 5 timing (GPU box only): one 500-taxon tree replicated to 64 chain states x 4,096 proposals each through
gsm_eval_dirty (host CSR arrays in pinned memory, result rows out).  usage: bench_proposals.py [chains] [proposals_per_chain] [taxa]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gammaspikemodel_b200 import GsmEngine, abi, synthetic


def pinned(x):
    t = torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    return t.numpy(), t


def run(chains=64, per_chain=4096, taxa=500, reps=5, device=0):
    b = synthetic.make_chain_states(chains, taxa)
    a = dict(height=b.dense("height"), parent=b.dense("parent"), rate=b.dense("rate"), spike=b.dense("spike"),
             nstubs=b.dense("nstubs"), params=b.params.numpy().copy(), use_spike=b.use_spike.numpy().copy())
    t0 = time.perf_counter()
    pr = synthetic.make_proposals(a, chains * per_chain)
    t_gen = time.perf_counter() - t0
    keep = []
    for k in list(pr):
        pr[k], t = pinned(pr[k])
        keep.append(t)
    P = chains * per_chain
    out, t_out = pinned(np.empty((P, abi.NOUT)))
    eng = GsmEngine(device=device)
    eng.set_mode(b.stub_mode, b.flags)
    eng.bind_batch(b.to("cuda:%d" % device))
    eng.eval_dirty(out=out, **pr)      # warm-up: allocations, base of the resident batch
    l0 = eng.launch_count
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        eng.eval_dirty(out=out, **pr)
        best = min(best, time.perf_counter() - t0)
    launches = (eng.launch_count - l0) // reps
    n_full = eng.last_dirty_full_count
    D = int(pr["dirty_ptr"][-1])
    h2d = sum(v.nbytes for v in pr.values())
    res = {"workload": "C5: one %d-taxon tree (%d nodes) x %d chain states x %d proposals" % (taxa, b.n_nodes, chains, per_chain),
           "proposals": P, "dirty_entries": D, "full_path_proposals": n_full, "incremental_proposals": P - n_full,
           "seconds": round(best, 5), "proposals_per_s": P / best, "dirty_entries_per_s": D / best,
           "equivalent_full_reevaluation_branch_evals_per_s": P * b.n_nodes / best,
           "h2d_bytes": h2d, "d2h_bytes": out.nbytes, "finite_frac": float(np.isfinite(out[:, 0]).mean()),
           "launches_per_call": launches, "generate_s": round(t_gen, 1)}
    eng.close()
    return res


if __name__ == "__main__":
    args = [int(x) for x in sys.argv[1:4]]
    print(json.dumps(run(*args)), flush=True)
