#!/usr/bin/env python
"""BASELINE config 5 timing (GPU box only): 64 chain states x 4,096 proposals each on 500-taxon trees through
gsm_eval_dirty (host CSR arrays in, result rows out).  usage: bench_proposals.py [chains] [proposals_per_chain] [taxa]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gammaspikemodel_b200 import GsmEngine, abi, synthetic

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 64
per_chain = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
taxa = int(sys.argv[3]) if len(sys.argv) > 3 else 500
b = synthetic.make_batch("C2", chains, taxa, seed=synthetic.SEEDS["C5"])
a = dict(height=b.dense("height"), parent=b.dense("parent"), rate=b.dense("rate"), spike=b.dense("spike"),
         nstubs=b.dense("nstubs"), params=b.params.numpy().copy(), use_spike=b.use_spike.numpy().copy())
t0 = time.perf_counter()
pr = synthetic.make_proposals(a, chains * per_chain)
t_gen = time.perf_counter() - t0
eng = GsmEngine()
eng.set_mode(b.stub_mode, b.flags)
eng.bind_batch(b.to("cuda"))
eng.eval_dirty(**{k: (v[:1024] if k not in ("dirty_ptr", "dirty_nodes", "new_height", "new_rate", "new_spike", "new_nstubs") else v)
                  for k, v in pr.items()} | {"dirty_ptr": pr["dirty_ptr"][:1025]} |
               {k: pr[k][: int(pr["dirty_ptr"][1024])] for k in ("dirty_nodes", "new_height", "new_rate", "new_spike", "new_nstubs")})   # warm-up
best = 1e9
for _ in range(3):
    t0 = time.perf_counter()
    out = eng.eval_dirty(**pr)
    best = min(best, time.perf_counter() - t0)
P = chains * per_chain
print(json.dumps({"workload": "C5: %d chains x %d proposals, %d taxa (%d nodes)" % (chains, per_chain, taxa, b.n_nodes),
                  "proposals": P, "dirty_entries": int(pr["dirty_ptr"][-1]), "seconds": round(best, 4),
                  "proposals_per_s": P / best, "branch_evals_per_s": P * b.n_nodes / best,
                  "finite_frac": float(np.isfinite(out[:, 0]).mean()), "launches": eng.launch_count,
                  "generate_s": round(t_gen, 1)}), flush=True)
eng.close()
