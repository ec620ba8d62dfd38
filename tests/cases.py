"""Case table shared by the CPU tier (kernel math emulated on the host vs oracle) and the GPU tier
(C ABI vs oracle): wirings, tree-prior variants, invalid states and edge shapes the reference handles."""
import numpy as np

from gammaspikemodel_b200 import abi, synthetic
from helpers import host_arrays, stubs_from_counts, variant

F = synthetic.DEFAULT_FLAGS
NOSTUBS = abi.F_HAS_INDICATOR | abi.F_HAS_RATES


def _base(kind, T, n, seed=None):
    b = synthetic.make_batch(kind, T, n, seed=seed)
    return host_arrays(b)


def swap_labels(a, t, i, j):
    """renumber nodes i <-> j of tree t (same tree, different Node.getNr())"""
    for k in ("height", "rate", "spike", "nstubs"):
        a[k][t, [i, j]] = a[k][t, [j, i]]
    p = a["parent"][t].copy()
    p[[i, j]] = p[[j, i]]
    q = p.copy()
    q[p == i] = j
    q[p == j] = i
    a["parent"][t] = q


def reference_tree_case():
    """the reference's own 36-taxon FBD tree with sampled ancestors (examples/stubs/nrStubs.xml:34-40; tests/golden)"""
    from test_reference_fixture import reference_tree
    return reference_tree(5)


def build_cases():
    """yields (name, arrays, flags, stub_mode, node_count)"""
    c2, c3, c4 = _base("C2", 24, 40), _base("C3", 24, 60), _base("C4", 24, 33)
    out = []
    add = lambda name, a, flags=F, mode=abi.STUBS_COUNTS, nc=None: out.append((name, a, flags, mode, nc))
    ref = reference_tree_case()
    add("reference-nrstubs-tree", ref)                                     # FBDWS, psi > 0, 10 sampled ancestors, 36 taxa
    add("reference-nrstubs-tree-nostub", ref, F, abi.STUBS_NOSTUB)         # nrStubs.xml / stochasticStubs.xml wirings
    add("reference-nrstubs-tree-cond-root", ref, F | abi.F_COND_ROOT | abi.F_COND_SAMPLING)
    # ---- wirings ----
    add("c2-default", c2)
    add("c3-default", c3)
    add("c4-default", c4)
    add("c3-no-stubs-wired", c3, NOSTUBS)                                  # PRCM:192 / BSP:81 stubs == null
    add("c3-stubs-only-tree-prior", c3, NOSTUBS | abi.F_STUBS_TO_TREE_PRIOR)
    add("c3-stubs-only-clock", c3, NOSTUBS | abi.F_STUBS_TO_CLOCK)
    add("c3-nostub-mixture", c3, F, abi.STUBS_NOSTUB)                      # BSP:102-172, quirk 2 of SURVEY §8a
    add("c2-nostub-mixture", c2, F, abi.STUBS_NOSTUB)
    add("c4-nostub-mixture", c4, F, abi.STUBS_NOSTUB)
    # ---- Stubs exact (reversible-jump) mode: stub lists, STP:573-600 / 639-663 ----
    def exact(a, seed):
        b = {k: v.copy() for k, v in a.items()}
        b["stubs"] = stubs_from_counts(b, seed)
        b["nstubs"] = np.zeros_like(b["nstubs"])      # not used in exact mode: the list defines the counts
        return b
    x3, x2, x4 = exact(c3, 1), exact(c2, 2), exact(c4, 3)
    add("c3-exact", x3, F, abi.STUBS_EXACT)                                # psi > 0: getLogQt(t, l, m, psi, rho)
    add("c2-exact", x2, F, abi.STUBS_EXACT)                                # psi == 0, rho == 1: getLogQt(t, l, m)
    add("c4-exact", x4, F, abi.STUBS_EXACT)
    add("c3-exact-ignore-stub-prior", x3, F | abi.F_IGNORE_STUB_PRIOR, abi.STUBS_EXACT)   # counts only: PRCM:195, BSP:81
    add("c3-exact-origin", variant(x3, F, origin=1e3), F | abi.F_ORIGIN, abi.STUBS_EXACT)
    b = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in x3.items()}
    ptr, br, tm = (v.copy() for v in b["stubs"])
    for t in range(0, len(ptr) - 1, 3):                                    # a stub outside its branch -> -inf (STP:580-582)
        if ptr[t + 1] > ptr[t]:
            tm[ptr[t]] = 1.5 if t % 2 else -0.25
    b["stubs"] = (ptr, br, tm)
    add("c3-exact-stub-outside-branch", b, F, abi.STUBS_EXACT)
    b = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in x3.items()}
    ptr, br, tm = (v.copy() for v in b["stubs"])
    n3x = b["height"].shape[1]
    for t in range(0, len(ptr) - 1, 2):                                    # a stub on the root: the Java throws (Stubs:631-634)
        if ptr[t + 1] > ptr[t]:
            br[ptr[t + 1] - 1] = n3x - 1
    b["stubs"] = (ptr, br, tm)
    add("c3-exact-stub-on-root", b, F, abi.STUBS_EXACT)
    b = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in x3.items()}
    ptr, br, tm = (v.copy() for v in b["stubs"])
    for t in range(len(ptr) - 1):                                          # a stub on a sampled ancestor (STP:232-240, 583-585)
        h, p = b["height"][t], b["parent"][t]
        sa = [i for i in range((n3x + 1) // 2) if p[i] >= 0 and h[p[i]] == h[i]]
        if sa and ptr[t + 1] > ptr[t] and t % 2 == 0:
            br[ptr[t]] = sa[0]
    b["stubs"] = (ptr, br, tm)
    add("c3-exact-stub-on-sampled-ancestor", b, F, abi.STUBS_EXACT)
    b = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in x2.items()}
    ptr, br, tm = (v.copy() for v in b["stubs"])
    br[::5] = 10_000                                                       # branch numbers that match no node: ignored
    br[1::7] = -3
    b["stubs"] = (ptr, br, tm)
    add("c2-exact-unmatched-branches", b, F, abi.STUBS_EXACT)
    b = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in x2.items()}
    nl2 = (b["height"].shape[1] + 1) // 2
    for t in range(0, b["height"].shape[0], 3):                            # zero-length internal branch carrying stubs:
        i = nl2 + 1                                                        # no h0 <= h1 rule in the RJ branch, E = 0
        b["height"][t, i] = b["height"][t, b["parent"][t, i]]
    add("c2-exact-zero-length-branches", b, F, abi.STUBS_EXACT)
    add("c2-exact-mu-zero", variant(x2, F, mu=0.0), F, abi.STUBS_EXACT)    # log(mu (e - 1)) = -inf (STP:826-831)
    add("c3-strict-clock", c3, F | abi.F_RELAXED_FALSE)                    # PRCM:216-220
    add("c3-no-rates", c3, F & ~abi.F_HAS_RATES)                           # PRCM:213
    add("c3-no-indicator", c3, F & ~abi.F_HAS_INDICATOR)
    add("c3-no-spike-on-dated-tips", c3, F | abi.F_NO_SPIKE_DATED_TIPS)    # PRCM:185-187
    add("c3-ignore-tree-prior", c3, F | abi.F_IGNORE_TREE_PRIOR)
    add("c3-ignore-stub-prior", c3, F | abi.F_IGNORE_STUB_PRIOR)
    # ---- tree-density variants (STP:179-214) ----
    big_origin = variant(c3, F, origin=1e3)
    add("c3-origin", big_origin, F | abi.F_ORIGIN)
    add("c3-origin-cond-sampling", big_origin, F | abi.F_ORIGIN | abi.F_COND_SAMPLING)
    add("c3-origin-cond-rho", big_origin, F | abi.F_ORIGIN | abi.F_COND_RHO_SAMPLING)
    add("c3-root-cond-sampling", c3, F | abi.F_COND_ROOT | abi.F_COND_SAMPLING)
    add("c3-root-cond-rho", c3, F | abi.F_COND_ROOT | abi.F_COND_RHO_SAMPLING)
    add("c3-origin-below-root", variant(c3, F, origin=1e-3), F | abi.F_ORIGIN)          # STP:182-184
    add("c3-origin-rho0", variant(big_origin, F, rho=0.0), F | abi.F_ORIGIN)            # STP:470 rho == 0 branch
    add("c2-psi0-rho-lt1", variant(c2, F, rho=0.37))                       # STP:351-379
    add("c3-psi0-rho1-dated", variant(c3, F, psi=0.0, rho=1.0))            # complete-rho on a dated-tip tree
    add("c3-psi0-rho-lt1-dated", variant(c3, F, psi=0.0))                  # STP:360-362 -> -inf (not all leaves extant)
    add("c2-psi-on-ultrametric", variant(c2, F, psi=0.3, rho=0.6))         # STP:290-347 with m = k = 0
    # ---- invalid / degenerate states (the reference returns -inf or NaN, never throws) ----
    add("c3-lambda-le-mu", variant(c3, F, lam=0.5, mu=0.7))                # STP:160-164
    add("c3-lambda-eq-mu", variant(c3, F, lam=0.5, mu=0.5))
    add("c3-shape-nonpositive", variant(c3, F, shape=0.0))                 # BSP:60-63
    add("c3-sigma-nonpositive", variant(c3, F, sigma=-1.0))                # BRP:36-39
    add("c3-psi-overflow", variant(c3, F, psi=800.0))                      # STP:176-178
    a = {k: v.copy() for k, v in c3.items()}
    a["spike"][3, 5] = -0.1                                                # BSP:72-75
    a["rate"][4, 7] = -2.0                                                 # BRP:47-50
    a["rate"][5, 2] = 0.0                                                  # BRP:168-170
    add("c3-negative-inputs", a)
    a = {k: v.copy() for k, v in c2.items()}
    a["spike"][0, 3] = 0.0                                                 # Gamma.logDensity(0): -inf / +inf / NaN by alpha
    a["params"][0, abi.P_SPIKE_SHAPE] = 3.0
    a["spike"][1, 3] = 0.0
    a["params"][1, abi.P_SPIKE_SHAPE] = 0.25
    a["nstubs"][1, 3] = 0
    a["spike"][2, 3] = 0.0
    a["params"][2, abi.P_SPIKE_SHAPE] = 1.0
    a["nstubs"][2, 3] = 0
    a["spike"][3, 4] = np.inf
    a["rate"][4, 4] = np.nan
    add("c2-zero-and-nonfinite-inputs", a)
    a = {k: v.copy() for k, v in c3.items()}
    for t in range(a["height"].shape[0]):                                  # SA leaf with stubs > 0 -> STP -inf (STP:232-240)
        h, p = a["height"][t], a["parent"][t]
        n_leaf = (a["height"].shape[1] + 1) // 2
        sa = [i for i in range(n_leaf) if p[i] >= 0 and h[p[i]] == h[i]]
        if sa and t % 2 == 0:
            a["nstubs"][t, sa[0]] = 2
    add("c3-sampled-ancestor-with-stubs", a)
    a = {k: v.copy() for k, v in c2.items()}
    n_leaf = (a["height"].shape[1] + 1) // 2
    for t in range(0, a["height"].shape[0], 3):                            # zero-length internal branch with / without stubs
        i = n_leaf + 1
        a["height"][t, i] = a["height"][t, a["parent"][t, i]]
        a["nstubs"][t, i] = 1 if t % 2 == 0 else 0                         # STP:558-561
    add("c2-zero-length-branches", a)
    a = variant(c2, F, lam=40.0, mu=1.0)                                   # exp(lambda*T) overflow -> NaN tree term (STP:420)
    a["height"] *= 30.0
    add("c2-exp-lambda-T-overflow", a)
    # sampled ancestors where the tree prior does not generate them (psi == 0): the lean kernel does not scan such trees
    # for fake parents, so every flavour of its bodies has to hand them to the generic path
    a = variant(c3, F, psi=0.0, rho=1.0)
    n3 = a["height"].shape[1]
    for t in range(a["height"].shape[0]):                                  # (a) root is not node n-1
        swap_labels(a, t, n3 - 1, n3 - 2 - (t % 5))
    add("c3-psi0-sa-root-not-last", a)
    odd = _base("C3", 24, 61)                                              # (b) L odd and the tail leaf L-1 is a sampled ancestor
    Lo = 61
    for t in range(odd["height"].shape[0]):
        odd["height"][t, Lo - 1] = odd["height"][t, odd["parent"][t, Lo - 1]]
        odd["spike"][t, Lo - 1] = 0.0
        odd["nstubs"][t, Lo - 1] = 0
    add("c3-psi0-sa-tail-leaf", variant(odd, F, psi=0.0, rho=1.0))
    add("c3-psi0-rho-lt1-sa-tail-leaf", variant(odd, F, psi=0.0, rho=0.4))
    add("c3-sa-tail-leaf", odd)                                            # same trees under psi > 0 (scanned)
    a = {k: v.copy() for k, v in c3.items()}
    for t in range(a["height"].shape[0]):
        swap_labels(a, t, n3 - 1, n3 - 2 - (t % 7))
    add("c3-root-not-last", a)                                             # GENERAL flavour with the scan
    # the Java's own overflow / domain behaviour outside the states BEAST can reach (reproduced, not fixed)
    a = {k: v.copy() for k, v in c3.items()}
    a["height"] *= 400.0                                                   # exp(c1 T) overflows: -inf, or NaN where an
    add("c3-exp-c1T-overflow", a)                                          # extinct leaf's own exp(c1 h) overflows too (STP:342)
    add("c3-exp-c1T-overflow-origin", variant(a, F, origin=1e5), F | abi.F_ORIGIN)
    for nm, src in (("c3", c3), ("c2", c2), ("c4", c4)):
        a = {k: v.copy() for k, v in src.items()}
        a["height"] -= 0.3                                                 # negative heights: log(mu (e^{(l-m)h} - 1)) is NaN (STP:826-831)
        add(nm + "-negative-heights", a)
    # ---- shapes ----
    add("two-taxa", _base("C3", 40, 2))
    add("single-node-trees", _base("C2", 7, 1))
    nc = np.array([(3 + 2 * (t % 15)) for t in range(24)], np.int32)       # ragged batch: odd node counts <= 33
    rag = {k: v.copy() for k, v in c4.items()}
    for t, n in enumerate(nc):                                             # re-root each tree on a prefix-closed subtree
        sub = synthetic.make_batch("C4", 1, (n + 1) // 2, seed=100 + t)
        for k in ("height", "parent", "rate", "spike", "nstubs"):
            rag[k][t, :n] = sub.dense(k)[0]
        rag["params"][t] = sub.params.numpy()[0]
    add("ragged-node-counts", rag, F, abi.STUBS_COUNTS, nc)
    return out


def empty_batch():
    return dict(height=np.zeros((0, 9)), parent=np.zeros((0, 9), np.int32), rate=np.zeros((0, 9)), spike=np.zeros((0, 9)),
                nstubs=np.zeros((0, 9), np.int32), params=np.zeros((0, abi.NPARAM)), use_spike=np.zeros((0,), np.uint8))
