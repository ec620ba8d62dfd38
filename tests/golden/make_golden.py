#!/usr/bin/env python
"""Regenerates tests/golden/golden_v1.npz: small seeded input batches with the oracle's outputs, plus the
reference's own JUnit goldens (test/gammaspike/distribution/StumpedTreePriorTest.java:35,53,71,96,114).

The reference is Java on BEAST 2.7.7 and cannot be run in this environment (no JVM), so these vectors come
from the CPU oracle (oracle/gsm_oracle.c), which is itself pinned by the JUnit goldens and the cross-checks of
tests/test_oracle_goldens.py.  They freeze today's agreed values so that later changes of either the oracle or
the kernels are caught.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from gammaspikemodel_b200 import abi, synthetic  # noqa: E402
from helpers import host_arrays, run_oracle  # noqa: E402

out = {}
for kind, T, n, mode in [("C2", 6, 12, abi.STUBS_COUNTS), ("C3", 6, 14, abi.STUBS_COUNTS), ("C4", 6, 10, abi.STUBS_COUNTS),
                         ("C3", 4, 9, abi.STUBS_NOSTUB)]:
    b = synthetic.make_batch(kind, T, n, seed=9000 + n)
    a = host_arrays(b)
    r = run_oracle(a, b.flags, mode, threads=1)
    key = "%s_%d_%d" % (kind, n, mode)
    for k, v in a.items():
        out[key + "/" + k] = v
    out[key + "/flags"] = np.array([b.flags, mode], np.int64)
    out[key + "/out"] = r["out"]
    out[key + "/rates"] = r["rates"]
    out[key + "/wspikes"] = r["wspikes"]
out["junit/height"] = np.array([[0.0, 1.0, 2.5, 3.5, 4.0, 2.0, 5.0]])
out["junit/parent"] = np.array([[5, 5, 4, 4, 6, 6, -1]], np.int32)
out["junit/expected"] = np.array([-28.7545, -28.3144, -30.6927, -17.9467, -20.1364])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), **out)
print("wrote golden_v1.npz with", len(out), "arrays")
