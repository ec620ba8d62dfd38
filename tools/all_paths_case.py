#!/usr/bin/env python
"""Small evaluations of every path (lean, generic, proposals) checked against the oracle in one short process (GPU box
only) — the case to put under a memory / race checker where one is available (compute-sanitizer is closed on this pool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from gammaspikemodel_b200 import GsmEngine, abi, synthetic
from helpers import host_arrays, run_capi, run_oracle, assert_results_match

for kind, T, n in (("C4", 12, 600), ("C3", 12, 300), ("C2", 24, 40), ("C4", 4, 2000)):
    b = synthetic.make_batch(kind, T, n)
    a = host_arrays(b)
    got = run_capi(a, b.flags, b.stub_mode)
    assert_results_match(got, run_oracle(a, b.flags, b.stub_mode), what=kind)
    got = run_capi(a, b.flags, abi.STUBS_NOSTUB)          # generic kernel (mixture mode)
    assert_results_match(got, run_oracle(a, b.flags, abi.STUBS_NOSTUB), what=kind + " nostub")
    pr = synthetic.make_proposals(a, 20)
    eng = GsmEngine(max_trees=T, max_nodes=b.n_nodes)
    run_capi(a, b.flags, b.stub_mode, engine=eng, rates=False)
    eng.eval_dirty(**pr)
    eng.close()
print("all_paths_case ok")
