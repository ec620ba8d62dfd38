#!/usr/bin/env python
"""A/B timing of one gsm_set_option on a resident batch (GPU box only). usage: tune_opt.py <kind> <trees> <taxa> <option> <v1> <v2> ..."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gammaspikemodel_b200 import _lib
if os.environ.get("GSM_LIB_PATH"):
    _lib.LIB_PATH = os.environ["GSM_LIB_PATH"]
from gammaspikemodel_b200 import GsmEngine, abi, synthetic
kind, trees, taxa, opt = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
vals = [int(v) for v in sys.argv[5:]]
dev = torch.device("cuda", 0)
b = synthetic.make_batch_chunked(kind, trees, taxa, device=dev, chunk=31250)
eng = GsmEngine(device=0); eng.set_mode(b.stub_mode, b.flags); eng.bind_batch(b)
out = torch.empty((b.n_trees, abi.NOUT), dtype=torch.float64, device=dev)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ref = None
for rep in range(int(os.environ.get("REPS", "2"))):
    for v in vals:
        eng.set_option(opt, v)
        for _ in range(3): eng.eval_device(out, None, None, st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): eng.eval_device(out, None, None, st.cuda_stream)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        cur = out.clone()
        same = True if ref is None else bool(torch.equal(torch.nan_to_num(cur), torch.nan_to_num(ref)))
        if ref is None: ref = cur
        print(json.dumps({"kind": kind, opt: v, "ms": round(ms, 4), "GBps": b.n_trees * b.n_nodes * 32 / ms / 1e6, "same_bits": same}), flush=True)
