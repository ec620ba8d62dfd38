#!/usr/bin/env python
"""Times gsm_eval_kernel launch configurations on one resident batch (GPU box only; tuning aid)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
from gammaspikemodel_b200 import _lib
if os.environ.get("GSM_LIB_PATH"):   # debug build with phase counters (tools/build_dbg.sh)
    _lib.LIB_PATH = os.environ["GSM_LIB_PATH"]
from gammaspikemodel_b200 import GsmEngine, abi, synthetic

kind, trees, taxa = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
# a configuration is BLOCKxMINB, optionally followed by "n" = no producer warp (128x3n), or "pk" = the persistent kernel
def parse(c):
    if c == "pk":
        return (0, 0, 1, 1)
    if c == "wk":                      # one warp per tree (small trees)
        return (0, 0, 0, -1)
    if c == "auto":
        return (0, 0, 0, 0)
    return tuple(map(int, c.rstrip("n").split("x"))) + (0 if c.endswith("n") else 1, 0)
cfgs = [parse(c) for c in sys.argv[4:]]
dev = torch.device("cuda", 0)
b = synthetic.make_batch_chunked(kind, trees, taxa, device=dev, chunk=31250)
if os.environ.get("GSM_ROOT_SWAP"):   # the root becomes node n-2 (trees as MCMC operators leave them)
    synthetic.move_root(b, b.n_nodes - 2)
eng = GsmEngine(device=0); eng.set_mode(b.stub_mode, b.flags); eng.bind_batch(b)
out = torch.empty((b.n_trees, abi.NOUT), dtype=torch.float64, device=dev)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ref = None
for blk, minb, prod, pk in cfgs:
    eng.set_option("block", blk); eng.set_option("minb", minb); eng.set_option("producer", prod); eng.set_option("persistent", max(pk, 0)); eng.set_option("warp_kernel", 1 if pk < 0 else 0)
    eng.set_option("smem_pad", int(os.environ.get("GSM_SMEM_PAD", "0")))
    for _ in range(3): eng.eval_device(out, None, None, st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): eng.eval_device(out, None, None, st.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    dbg = getattr(_lib.load(), "gsm_debug_phase_cycles", None) if os.environ.get("GSM_LIB_PATH") else None
    if dbg is not None:
        buf = (ctypes.c_ulonglong * 8)()
        dbg(buf, 1)
        for _ in range(2): eng.eval_device(out, None, None, st.cuda_stream)
        torch.cuda.synchronize()
        dbg(buf, 1)
        tot = sum(buf) or 1
        names = ["setup+requests", "wait rows", "repack", "pass A", "wait A/tables", "pass B", "wait B", "(in pass B: waiting for chunks)"]
        if pk: names = ["wait front end", "set-up", "pass A", "wait A / scan", "pass B", "combine+reduce", "wait B", "(in pass B: waiting for chunks; row is folded into the next wait)"]
        wdbg = getattr(_lib.load(), "gsm_debug_warp_cycles", None)
        if wdbg is not None:
            wb = (ctypes.c_ulonglong * 16)()
            wdbg(wb, 1)
            print("pass B cycles per tree by warp (chunk loops / + tail + combine): " + " ".join("%.0f/%.0f" % (wb[i] / (2 * b.n_trees), wb[i + 8] / (2 * b.n_trees)) for i in range(8)), flush=True)
        print("phase cycles per tree (thread 0): " + ", ".join("%s %.0f (%.0f%%)" % (n, c / (2 * b.n_trees), 100 * c / tot) for n, c in zip(names, buf)), flush=True)
    cur = out.clone()
    same = True if ref is None else bool(torch.equal(torch.nan_to_num(cur), torch.nan_to_num(ref)))
    maxdiff = 0.0 if ref is None else float((torch.nan_to_num(cur) - torch.nan_to_num(ref)).abs().max())
    if ref is None: ref = cur
    br = b.n_trees * b.n_nodes
    print(json.dumps({"kind": kind, "trees": trees, "taxa": taxa, "block": blk, "minb": minb, "producer": prod, "persistent": pk, "ms": round(ms, 4),
                      "branch_evals_per_s": br / ms * 1e3, "GBps": br * 32 / ms / 1e6, "bitwise_same_as_first": same, "max_abs_diff_vs_first": maxdiff}), flush=True)
