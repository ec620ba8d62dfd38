#!/usr/bin/env python
"""Times the wirings that do not take the lean path (GPU box only): mixture-mode spike prior (Stubs nostub mode, BSP:102-172),
noSpikeOnDatedTips, rates not wired.  usage: time_modes.py [trees] [taxa]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gammaspikemodel_b200 import GsmEngine, abi, synthetic

trees = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
taxa = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device("cuda", 0)
b = synthetic.make_batch_chunked("C3", trees, taxa, device=dev, chunk=16384)
F = b.flags
modes = [("default (lean path)", abi.STUBS_COUNTS, F),
         ("mixture spike prior (nostub)", abi.STUBS_NOSTUB, F),
         ("noSpikeOnDatedTips", abi.STUBS_COUNTS, F | abi.F_NO_SPIKE_DATED_TIPS),
         ("rates not wired", abi.STUBS_COUNTS, F & ~abi.F_HAS_RATES)]
out = torch.empty((b.n_trees, abi.NOUT), dtype=torch.float64, device=dev)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for name, mode, flags in modes:
    eng = GsmEngine(device=0); eng.set_mode(mode, flags); eng.bind_batch(b)
    for _ in range(2): eng.eval_device(out, None, None, st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): eng.eval_device(out, None, None, st.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    br = b.n_trees * b.n_nodes
    print(json.dumps({"mode": name, "trees": trees, "taxa": taxa, "ms": round(ms, 4), "branch_evals_per_s": br / ms * 1e3,
                      "finite_frac": float(torch.isfinite(out[:, 0]).double().mean())}), flush=True)
    eng.close()
