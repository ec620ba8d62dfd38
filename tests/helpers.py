"""Shared helpers of the test-suite: oracle / emulator / C-ABI runners on a synthetic.Batch and comparisons."""
import ctypes as C
import os

import numpy as np

import oracle
from gammaspikemodel_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-10   # north_star tolerance for fp64 log-densities, rates and ProportionOfSaltation
ATOL = 1e-9    # floor for sums of O(1e3) terms that happen to land near zero

_emu = None


def emu():
    global _emu
    if _emu is None:
        _emu = C.CDLL(os.path.join(ROOT, "tests", "_emu", "libgsm_emu.so"))
        _emu.gsm_emu_eval_batch.argtypes = [C.c_uint32, C.c_int32, C.c_int64, C.c_int32, C.c_int64] + [C.c_void_p] * 11
        _emu.gsm_emu_eval_batch.restype = C.c_int
        _emu.gsm_emu_eval_batch_lean.argtypes = _emu.gsm_emu_eval_batch.argtypes + [C.c_void_p]
        _emu.gsm_emu_eval_batch_lean.restype = C.c_int
        _emu.gsm_emu_eval_batch_stubs.argtypes = [C.c_uint32, C.c_int32, C.c_int64, C.c_int32, C.c_int64] + [C.c_void_p] * 14
        _emu.gsm_emu_eval_batch_stubs.restype = C.c_int
    return _emu


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def host_arrays(b):
    """dense [T][N] numpy arrays of a synthetic.Batch (the layout the C ABI takes)"""
    return dict(height=b.dense("height"), parent=b.dense("parent"), rate=b.dense("rate"), spike=b.dense("spike"),
                nstubs=b.dense("nstubs"), params=b.params.cpu().numpy().copy(), use_spike=b.use_spike.cpu().numpy().copy())


def run_oracle(a, flags, mode, node_count=None, rates=True, threads=0):
    r = oracle.eval_batch(flags, mode, a["height"], a["parent"], a["rate"], a["spike"], a["nstubs"], a["params"],
                          a["use_spike"], node_count=node_count, want_rates=rates, want_wspikes=rates, n_threads=threads,
                          stubs=a.get("stubs"))
    return r


def stubs_from_counts(a, seed=0, n_nodes=None):
    """EXACT-mode stub lists (CSR by tree) that realise the per-branch counts of a["nstubs"] (entries 0 .. n-2; the root
    carries no stub): uniform relative times along each branch."""
    rng = np.random.default_rng(seed)
    T, N = a["nstubs"].shape
    cnt = a["nstubs"].copy()
    cnt[:, N - 1:] = 0
    if n_nodes is not None:
        for t in range(T):
            cnt[t, n_nodes[t] - 1:] = 0
    cnt = np.maximum(cnt, 0)
    ptr = np.zeros(T + 1, np.int64)
    ptr[1:] = np.cumsum(cnt.sum(axis=1))
    branch = np.repeat(np.tile(np.arange(N, dtype=np.int32), T), cnt.ravel())
    time = rng.random(branch.shape[0])
    # shuffle inside each tree: the list is unordered in the reference
    for t in range(T):
        perm = rng.permutation(ptr[t + 1] - ptr[t]) + ptr[t]
        branch[ptr[t]:ptr[t + 1]] = branch[perm]
    return ptr, branch.astype(np.int32), time


def run_emu(a, flags, mode, node_count=None, rates=True):
    T, N = a["height"].shape
    out = np.empty((T, abi.NOUT))
    r = np.full((T, N), np.nan) if rates else None
    w = np.full((T, N), np.nan) if rates else None
    nc = None if node_count is None else np.ascontiguousarray(node_count, np.int32)
    if a.get("stubs") is not None:
        sp, sb, st = (np.ascontiguousarray(x, d) for x, d in zip(a["stubs"], (np.int64, np.int32, np.float64)))
        r = np.full((T, N), np.nan)
        w = np.full((T, N), np.nan)
        emu().gsm_emu_eval_batch_stubs(flags, mode, T, N, N, _p(a["height"]), _p(a["parent"]), _p(a["rate"]), _p(a["spike"]),
                                       _p(a["nstubs"]), _p(a["params"]), _p(a["use_spike"]), _p(nc), _p(sp), _p(sb), _p(st),
                                       _p(out), _p(r), _p(w))
        return {"out": out, "rates": r if rates else None, "wspikes": w if rates else None}
    emu().gsm_emu_eval_batch(flags, mode, T, N, N, _p(a["height"]), _p(a["parent"]), _p(a["rate"]), _p(a["spike"]),
                             _p(a["nstubs"]), _p(a["params"]), _p(a["use_spike"]), _p(nc), _p(out), _p(r), _p(w))
    return {"out": out, "rates": r, "wspikes": w}


def run_emu_lean(a, flags, mode, node_count=None, rates=True):
    """host build of the lean path (gsm_lean_kernel + gsm_post_kernel; trees it hands back go through the generic walk, as on
    the device).  Returns None when the wiring is outside its family; else the results plus "n_lean" = trees it finished."""
    T, N = a["height"].shape
    out = np.empty((T, abi.NOUT))
    r = np.full((T, N), np.nan) if rates else None
    w = np.full((T, N), np.nan) if rates else None
    nc = None if node_count is None else np.ascontiguousarray(node_count, np.int32)
    n_lean = np.zeros(1, np.int64)
    if a.get("stubs") is not None and not (flags & abi.F_IGNORE_STUB_PRIOR) and (flags & abi.F_STUBS_TO_TREE_PRIOR):
        return None   # reversible-jump placement terms: generic path only (gsm_lean_wiring_ok)
    if a.get("stubs") is not None:
        return None   # the emulated lean walk takes counts, not lists
    rc = emu().gsm_emu_eval_batch_lean(flags, mode, T, N, N, _p(a["height"]), _p(a["parent"]), _p(a["rate"]), _p(a["spike"]),
                                       _p(a["nstubs"]), _p(a["params"]), _p(a["use_spike"]), _p(nc), _p(out), _p(r), _p(w),
                                       _p(n_lean))
    if rc != 0:
        return None
    return {"out": out, "rates": r, "wspikes": w, "n_lean": int(n_lean[0])}


def run_capi(a, flags, mode, node_count=None, rates=True, engine=None):
    from gammaspikemodel_b200 import GsmEngine
    T, N = a["height"].shape
    eng = engine or GsmEngine(max_trees=max(T, 1), max_nodes=max(N, 1))
    eng.set_mode(mode, flags)
    eng.upload_trees(a["height"], a["parent"], node_count)
    eng.upload_branch_params(a["rate"], a["spike"], a["nstubs"])
    eng.upload_tree_params(a["params"], a["use_spike"])
    if a.get("stubs") is not None:
        eng.upload_stubs(*a["stubs"])
    r = eng.eval(want_rates=rates, want_wspikes=rates)
    if engine is None:
        eng.close()
    return r


def rel_err(got, ref):
    got, ref = np.asarray(got, float), np.asarray(ref, float)
    with np.errstate(all="ignore"):
        d = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
    same = (got == ref) | (np.isnan(got) & np.isnan(ref))
    return np.where(same, 0.0, d)


def assert_close(got, ref, rtol=RTOL, atol=ATOL, what=""):
    """got == ref within rtol (relative) or atol (absolute); NaN matches NaN, inf matches the same inf."""
    got, ref = np.asarray(got, float), np.asarray(ref, float)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    same = (got == ref) | (np.isnan(got) & np.isnan(ref))
    with np.errstate(all="ignore"):
        ok = same | (np.abs(got - ref) <= atol + rtol * np.abs(ref))
    if not np.all(ok):
        bad = np.argwhere(~ok)
        i = tuple(bad[0])
        raise AssertionError("%s: %d mismatches; first at %s: got %r, ref %r (rel %.3g)"
                             % (what, len(bad), i, got[i], ref[i], rel_err(got[i], ref[i])))


def assert_results_match(got, ref, n_nodes=None, node_count=None, what=""):
    assert_close(got["out"], ref["out"], what=what + " out_tree")
    # integer-valued output: bit-exact
    assert np.array_equal(got["out"][:, abi.O_N_EXTANT], ref["out"][:, abi.O_N_EXTANT]), what + " nExtant"
    for key in ("rates", "wspikes"):
        if got.get(key) is not None and ref.get(key) is not None:
            g, r = got[key], ref[key]
            if node_count is not None:
                mask = np.arange(g.shape[1])[None, :] < np.asarray(node_count)[:, None]
                g, r = np.where(mask, g, 0.0), np.where(mask, r, 0.0)
            assert_close(g, r, atol=0.0, what=what + " " + key)


def variant(a, flags, **edits):
    """copy of the host arrays with per-tree parameter edits, e.g. psi=0.0, rho=1.0"""
    b = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in a.items()}
    idx = dict(clock_rate=abi.P_CLOCK_RATE, spike_mean=abi.P_SPIKE_MEAN, shape=abi.P_SPIKE_SHAPE, sigma=abi.P_CLOCK_SD,
               lam=abi.P_LAMBDA, mu=abi.P_MU, psi=abi.P_PSI, rho=abi.P_RHO, origin=abi.P_ORIGIN)
    for k, v in edits.items():
        b["params"][:, idx[k]] = v
    return b


def apply_proposals(a, pr):
    """host arrays of the proposed states [P][N] (what gsm_eval_dirty evaluates): tree tree_of_prop[p] with the CSR edits"""
    t = pr["tree_of_prop"]
    b = {k: a[k][t].copy() for k in ("height", "parent", "rate", "spike", "nstubs")}
    b["params"] = pr["new_params"].copy() if pr.get("new_params") is not None else a["params"][t].copy()
    b["use_spike"] = pr["new_use_spike"].copy() if pr.get("new_use_spike") is not None else a["use_spike"][t].copy()
    if pr.get("height_scale") is not None:                # BEAST Tree.scale: internal nodes that are not fake
        N = b["height"].shape[1]
        for q in np.nonzero(pr["height_scale"] != 1.0)[0]:
            h, p = b["height"][q], b["parent"][q]
            internal = np.zeros(N, bool)
            internal[p[p >= 0]] = True
            leaf_da = ~internal & (p >= 0) & (h == h[np.maximum(p, 0)])
            fake = np.zeros(N, bool)
            fake[p[leaf_da]] = True
            sel = internal & ~fake
            b["height"][q, sel] = h[sel] * pr["height_scale"][q]
    ptr = pr["dirty_ptr"]
    rows = np.repeat(np.arange(len(t)), np.diff(ptr))
    cols = pr["dirty_nodes"]
    for key, name in (("height", "new_height"), ("rate", "new_rate"), ("spike", "new_spike"), ("nstubs", "new_nstubs")):
        if pr.get(name) is not None:
            b[key][rows, cols] = pr[name]
    return b


def stub_term_noise(a):
    """Rounding noise of the REFERENCE's own stub density per tree (psi == 0, rho == 1 states): the Java evaluates
    E = I(h1) - I(h0) with I(h) = 2 (log(l e^{(l-m)h} - m) + l (T - h)) (STP:758-769), a difference of O(l T) numbers, so a
    branch with a tiny expected stub count E and n > 0 stubs contributes a few n * ulp(I) / E of noise to n log E.  Used where
    randomly proposed states hit such branches: the oracle (same operation order as the Java) is then itself only accurate
    to this bound, and 50-digit arithmetic sides with the device value (tests/test_conditioning.py)."""
    h, par, ns, prm = a["height"], a["parent"], a["nstubs"], a["params"]
    lam, mu = prm[:, abi.P_LAMBDA:abi.P_LAMBDA + 1], prm[:, abi.P_MU:abi.P_MU + 1]
    T = h.max(axis=1, keepdims=True)
    hp = np.take_along_axis(h, np.maximum(par, 0), axis=1)
    with np.errstate(all="ignore"):
        I = lambda x: 2 * (np.log(lam * np.exp((lam - mu) * x) - mu) + lam * (T - x))
        E = I(h) - I(hp)
        mag = np.maximum(np.abs(I(h)), np.abs(I(hp)))
        noise = np.where((par >= 0) & (ns > 0) & (hp > h), ns * 32 * np.spacing(mag) / np.maximum(E, 1e-300), 0.0)
    return noise.sum(axis=1)
