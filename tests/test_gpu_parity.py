"""-m gpu: the sm_100a path, called through the C ABI (libgsm_b200.so), against the CPU oracle on the same
seeded inputs.  Tolerances: 1e-10 relative for fp64 log-densities / rates / ProportionOfSaltation (north_star),
bit-exact for integer outputs (nExtant) and for values that are pure selections (weighted spikes)."""
import numpy as np
import pytest

import oracle
from gammaspikemodel_b200 import GsmEngine, GsmError, abi, synthetic
from cases import build_cases, empty_batch
from helpers import assert_close, assert_results_match, host_arrays, run_capi, run_oracle

pytestmark = pytest.mark.gpu

F = synthetic.DEFAULT_FLAGS
CASES = build_cases()


@pytest.mark.parametrize("kind,T,n", [("C2", 256, 200), ("C3", 128, 200), ("C4", 64, 500), ("C3", 32, 1000), ("C4", 8, 2000),
                                      ("C2", 300, 5), ("C3", 100, 2), ("C2", 50, 1), ("C3", 3, 4000)])
def test_default_wiring_matches_oracle(kind, T, n):
    b = synthetic.make_batch(kind, T, n)
    a = host_arrays(b)
    ref = run_oracle(a, b.flags, b.stub_mode)
    got = run_capi(a, b.flags, b.stub_mode)
    assert_results_match(got, ref, what="%s %dx%d" % (kind, T, n))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_case_table_matches_oracle(case):
    name, a, flags, mode, nc = case
    if name in ("c3-lambda-le-mu", "c3-lambda-eq-mu") and mode == abi.STUBS_NOSTUB:
        pytest.skip("mixture-mode spike prior is not defined for lambda <= mu")
    ref = run_oracle(a, flags, mode, node_count=nc)
    got = run_capi(a, flags, mode, node_count=nc)
    assert_results_match(got, ref, node_count=nc, what=name)


def test_logp_only_launch_matches_oracle():
    b = synthetic.make_batch("C3", 64, 300)
    a = host_arrays(b)
    ref = run_oracle(a, b.flags, b.stub_mode)
    got = run_capi(a, b.flags, b.stub_mode, rates=False)
    assert got["rates"] is None
    assert_results_match(got, ref, what="logp-only")


def test_empty_batch_and_argument_errors():
    a = empty_batch()
    eng = GsmEngine(max_trees=4, max_nodes=9)
    eng.set_mode(abi.STUBS_COUNTS, F)
    eng.upload_trees(a["height"], a["parent"])
    eng.upload_branch_params(a["rate"], a["spike"], a["nstubs"])
    eng.upload_tree_params(a["params"], a["use_spike"])
    assert eng.eval()["out"].shape == (0, abi.NOUT)
    with pytest.raises(GsmError):                       # exceeds the handle capacity
        eng.upload_trees(np.zeros((5, 9)), np.zeros((5, 9), np.int32))
    with pytest.raises(GsmError):                       # conditionOnSampling + conditionOnRhoSampling (STP:92-94)
        eng.set_mode(abi.STUBS_COUNTS, abi.F_COND_SAMPLING | abi.F_COND_RHO_SAMPLING)
    with pytest.raises(GsmError):                       # too many nodes for the shared-memory kernels
        GsmEngine(max_trees=1, max_nodes=abi.MAX_NODES + 1)
    eng2 = GsmEngine(max_trees=2, max_nodes=9)
    with pytest.raises(GsmError):                       # eval before upload
        eng2.eval()
    eng.close()
    eng2.close()


def test_state_updates_between_evals():
    """MCMC-style use of one handle: re-upload only what changed (rates / spikes / per-tree scalars)."""
    b = synthetic.make_batch("C3", 32, 100)
    a = host_arrays(b)
    eng = GsmEngine(max_trees=32, max_nodes=b.n_nodes)
    first = run_capi(a, b.flags, b.stub_mode, engine=eng)
    a2 = {k: v.copy() for k, v in a.items()}
    a2["rate"] *= 1.3
    a2["params"][:, abi.P_CLOCK_SD] *= 0.9
    a2["use_spike"][:] = 1 - a2["use_spike"]
    eng.upload_branch_params(rate=a2["rate"])
    eng.upload_tree_params(a2["params"], a2["use_spike"])
    second = eng.eval(want_rates=True, want_wspikes=True)
    assert_results_match(second, run_oracle(a2, b.flags, b.stub_mode), what="after update")
    assert not np.array_equal(first["out"], second["out"])
    eng.close()


@pytest.mark.parametrize("kind,T,n", [("C3", 700, 150), ("C4", 37, 1200)])
def test_streaming_host_path_equals_resident_path(kind, T, n):
    """gsm_eval_host (chunked, double-buffered) must return bit-identical rows to upload + gsm_eval."""
    b = synthetic.make_batch(kind, T, n)
    a = host_arrays(b)
    eng = GsmEngine(max_trees=T, max_nodes=b.n_nodes)
    eng.set_option("host_chunk_mb", 1)                   # force many chunks
    resident = run_capi(a, b.flags, b.stub_mode, engine=eng)
    eng.set_mode(b.stub_mode, b.flags)
    streamed = eng.eval_host(a["height"], a["parent"], a["spike"], a["params"], rate=a["rate"], nstubs=a["nstubs"],
                             use_spike=a["use_spike"], want_rates=True, want_wspikes=True)
    for k in ("out", "rates", "wspikes"):
        assert np.array_equal(streamed[k], resident[k], equal_nan=True), k
    eng.close()


def test_device_resident_path_equals_host_path():
    import torch
    b = synthetic.make_batch("C3", 200, 120)
    a = host_arrays(b)
    host = run_capi(a, b.flags, b.stub_mode)
    d = b.to("cuda")
    eng = GsmEngine()
    eng.set_mode(b.stub_mode, b.flags)
    eng.bind_batch(d)
    out = torch.empty((b.n_trees, abi.NOUT), dtype=torch.float64, device="cuda")
    rates = torch.full((b.n_trees, b.stride), float("nan"), dtype=torch.float64, device="cuda")
    st = torch.cuda.Stream()
    eng.eval_device(out, rates, None, st.cuda_stream)
    st.synchronize()
    assert np.array_equal(out.cpu().numpy(), host["out"], equal_nan=True)
    assert np.array_equal(rates[:, : b.n_nodes].cpu().numpy(), host["rates"])
    assert eng.launch_count in (1, 2)   # generic: one kernel; lean path: lean + post kernel
    eng.close()


def test_stub_cdf_matches_oracle():
    """BranchSpikePrior.getCumulativeProbs (BSP:315-389) with mu from StumpedTreePrior.getMeanStubNumber."""
    b = synthetic.make_batch("C3", 8, 40)
    a = host_arrays(b)
    flags = b.flags | abi.F_BSP_HAS_INDICATOR
    eng = GsmEngine(max_trees=8, max_nodes=b.n_nodes)
    run_capi(a, flags, abi.STUBS_NOSTUB, engine=eng)
    L = oracle.lib()
    checked = 0
    for t in range(8):
        lam, mu_, psi, rho = a["params"][t, abi.P_LAMBDA:abi.P_RHO + 1]
        T_root = a["height"][t].max()
        for node in range(0, b.n_nodes - 1, 7):
            h, hp = a["height"][t, node], a["height"][t, a["parent"][t, node]]
            mean = L.gsmo_mean_stub_number(h, hp, lam, mu_, psi, rho, T_root)
            if not mean > 0:
                continue
            ref, n_ref = oracle.cumulative_probs(mean, a["spike"][t, node], a["params"][t, abi.P_SPIKE_SHAPE], a["use_spike"][t])
            got, n_got = eng.stub_cdf(t, node)
            if a["spike"][t, node] == 0 and a["use_spike"][t]:
                continue                                  # all-zero weights: 0/0 in the Java as well
            assert n_got == n_ref
            assert_close(got, ref, rtol=1e-9, atol=1e-12, what="cdf t=%d node=%d" % (t, node))
            checked += 1
    assert checked > 20
    eng.close()


def test_full_size_properties_c2():
    """BASELINE config 2 at full size (10k x 200 taxa): size-independent checks instead of a full oracle run:
    tree-permutation equivariance, batch-split invariance, a random 2% sample against the oracle."""
    import torch
    b = synthetic.make_batch("C2", 10_000, 200)
    a = host_arrays(b)
    full = run_capi(a, b.flags, b.stub_mode, rates=False)["out"]
    rng = np.random.default_rng(0)
    perm = rng.permutation(b.n_trees)
    ap = {k: v[perm] for k, v in a.items()}
    assert np.array_equal(run_capi(ap, b.flags, b.stub_mode, rates=False)["out"], full[perm], equal_nan=True)
    half = {k: v[:5000] for k, v in a.items()}
    assert np.array_equal(run_capi(half, b.flags, b.stub_mode, rates=False)["out"], full[:5000], equal_nan=True)
    idx = np.sort(rng.choice(b.n_trees, 200, replace=False))
    sub = {k: v[idx] for k, v in a.items()}
    assert_close(full[idx], run_oracle(sub, b.flags, b.stub_mode, rates=False)["out"], what="C2 sample")
    assert np.array_equal(full[:, abi.O_N_EXTANT], np.full(b.n_trees, 200.0))


def test_cpp_host_mirror_runs_the_reference_junit_test():
    """gammaspikemodel_b200/host/stumped_tree_prior_test.cpp = StumpedTreePriorTest.java through the C++ host
    mirror and the C ABI."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gammaspikemodel_b200", "host",
                       "stumped_tree_prior_test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all passed" in r.stdout


@pytest.mark.parametrize("kind,T,n", [("C4", 8192, 2000), ("C3", 8192, 1000), ("C2", 10_000, 200)])
def test_every_tree_of_a_large_batch_matches_oracle(kind, T, n):
    """Every tree of a batch large enough to keep all SMs busy for many waves, device-resident path and streaming
    host path, against the oracle (not a sample: an intra-CTA ordering bug shows up in ~1% of the trees only)."""
    import torch
    b = synthetic.make_batch_chunked(kind, T, n, seed=synthetic.SEEDS[kind], device="cuda", chunk=2048)
    N = b.n_nodes
    a = dict(height=b.height[:, :N].contiguous().cpu().numpy(), parent=b.parent[:, :N].contiguous().cpu().numpy(),
             rate=b.rate[:, :N].contiguous().cpu().numpy(), spike=b.spike[:, :N].contiguous().cpu().numpy(),
             nstubs=b.nstubs[:, :N].contiguous().cpu().numpy(), params=b.params.cpu().numpy(),
             use_spike=b.use_spike.cpu().numpy())
    ref = run_oracle(a, b.flags, b.stub_mode, rates=False, threads=0)["out"]
    eng = GsmEngine()
    eng.set_mode(b.stub_mode, b.flags)
    eng.bind_batch(b)
    out = torch.empty((T, abi.NOUT), dtype=torch.float64, device="cuda")
    st = torch.cuda.Stream()
    for _ in range(3):                                   # repeated launches: same bits every time
        eng.eval_device(out, None, None, st.cuda_stream)
        st.synchronize()
        dev = out.cpu().numpy()
        assert_close(dev, ref, what="%s device-resident, all trees" % kind)
    streamed = eng.eval_host(a["height"], a["parent"], a["spike"], a["params"], rate=a["rate"], nstubs=a["nstubs"],
                             use_spike=a["use_spike"])["out"]
    assert np.array_equal(streamed, dev, equal_nan=True)
    eng.close()


@pytest.mark.parametrize("kind", ["C4", "C3"])
def test_ragged_batch_mixed_tree_sizes_matches_oracle(kind):
    """One launch over trees of very different sizes (node_count per tree): the launch configuration is picked for the
    largest tree, the chunk / tail logic of the lean kernel sees 3 ... 1,999 nodes."""
    sizes = [2, 3, 26, 129, 300, 513, 777, 1000]
    per = 96
    Nmax = 2 * max(sizes) - 1
    parts, counts = [], []
    for k, taxa in enumerate(sizes):
        b = synthetic.make_batch(kind, per, taxa, seed=synthetic.SEEDS[kind] + k)
        a = host_arrays(b)
        n = b.n_nodes
        pad = {}
        for key, fill in (("height", 0.0), ("parent", -1), ("rate", 1.0), ("spike", 0.0), ("nstubs", 0)):
            x = np.full((per, Nmax), fill, dtype=a[key].dtype)
            x[:, :n] = a[key]
            pad[key] = x
        pad["params"], pad["use_spike"] = a["params"], a["use_spike"]
        parts.append(pad)
        counts.append(np.full(per, n, np.int32))
    a = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    nc = np.concatenate(counts)
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(nc))                      # neighbours in the grid have different sizes
    a = {k: np.ascontiguousarray(v[perm]) for k, v in a.items()}
    nc = np.ascontiguousarray(nc[perm])
    F_ = synthetic.DEFAULT_FLAGS
    ref = run_oracle(a, F_, abi.STUBS_COUNTS, node_count=nc)
    got = run_capi(a, F_, abi.STUBS_COUNTS, node_count=nc)
    assert_results_match(got, ref, node_count=nc, what="ragged " + kind)


def test_exact_mode_needs_the_stub_list_where_the_reference_uses_it():
    """GSM_STUBS_EXACT: the point-process stub density (STP:573-600 / 639-663) is defined by the stub list; without it the
    call fails instead of returning the count-mode density (round-1 parity hole)."""
    from helpers import stubs_from_counts
    b = synthetic.make_batch("C3", 16, 40)
    a = host_arrays(b)
    eng = GsmEngine(max_trees=16, max_nodes=b.n_nodes)
    eng.set_mode(abi.STUBS_EXACT, F)
    eng.upload_trees(a["height"], a["parent"])
    eng.upload_branch_params(a["rate"], a["spike"], a["nstubs"])
    eng.upload_tree_params(a["params"], a["use_spike"])
    with pytest.raises(GsmError) as ei:
        eng.eval()
    assert ei.value.code == abi.E_STATE
    with pytest.raises(GsmError) as ei:
        eng.eval_host(a["height"], a["parent"], a["spike"], a["params"], rate=a["rate"], nstubs=a["nstubs"], use_spike=a["use_spike"])
    assert ei.value.code == abi.E_UNSUPPORTED
    with pytest.raises(GsmError) as ei:
        eng.eval_dirty(np.array([0], np.int32), np.array([0, 0], np.int64), np.zeros(0, np.int32))
    assert ei.value.code == abi.E_UNSUPPORTED
    # stub prior not evaluated: the per-branch counts are all PRCM / BSP need, with or without the list
    eng.set_mode(abi.STUBS_EXACT, F | abi.F_IGNORE_STUB_PRIOR)
    by_counts = eng.eval(want_rates=True, want_wspikes=True)
    ax = dict(a)
    ax["stubs"] = stubs_from_counts(a, 3)
    ref = run_oracle(ax, F | abi.F_IGNORE_STUB_PRIOR, abi.STUBS_EXACT)
    assert_results_match(by_counts, ref, what="exact mode, caller-supplied counts")
    eng.upload_stubs(*ax["stubs"])
    by_list = eng.eval(want_rates=True, want_wspikes=True)
    for k in ("out", "rates", "wspikes"):
        assert np.array_equal(by_counts[k], by_list[k], equal_nan=True), k
    # ... and with the list the full density is available
    eng.set_mode(abi.STUBS_EXACT, F)
    assert_results_match(eng.eval(want_rates=True, want_wspikes=True), run_oracle(ax, F, abi.STUBS_EXACT), what="exact mode")
    eng.close()


def test_exact_mode_device_bound_lists():
    import torch
    from helpers import stubs_from_counts
    b = synthetic.make_batch("C3", 300, 150)
    a = host_arrays(b)
    a["stubs"] = stubs_from_counts(a, 9)
    ref = run_oracle(a, F, abi.STUBS_EXACT, rates=False)["out"]
    d = b.to("cuda")
    sp, sb, st = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in a["stubs"])
    eng = GsmEngine()
    eng.set_mode(abi.STUBS_EXACT, F)
    eng.bind_batch(d)
    eng.bind_stubs(sp, sb, st)
    out = torch.empty((b.n_trees, abi.NOUT), dtype=torch.float64, device="cuda")
    eng.eval_device(out)
    eng.sync()
    assert_close(out.cpu().numpy(), ref, what="exact mode, device-bound")
    eng.close()


def test_two_handles_on_two_threads_give_identical_rows():
    """one handle per chain, handles usable from any thread, no process-global state.  Two handles with trees of
    different sizes (different launch configurations and shared-memory sizes) evaluate concurrently, many times."""
    import threading
    ba, bb = synthetic.make_batch("C4", 600, 700, seed=5), synthetic.make_batch("C3", 900, 90, seed=6)
    aa, ab = host_arrays(ba), host_arrays(bb)
    ref_a, ref_b = run_capi(aa, ba.flags, ba.stub_mode), run_capi(ab, bb.flags, bb.stub_mode)
    errors = []

    def work(a, b, ref, reps):
        try:
            eng = GsmEngine(max_trees=b.n_trees, max_nodes=b.n_nodes)
            for _ in range(reps):
                got = run_capi(a, b.flags, b.stub_mode, engine=eng)
                for k in ("out", "rates", "wspikes"):
                    if not np.array_equal(got[k], ref[k], equal_nan=True):
                        errors.append(k)
            eng.close()
        except Exception as e:   # noqa: BLE001
            errors.append(repr(e))

    th = [threading.Thread(target=work, args=(aa, ba, ref_a, 12)), threading.Thread(target=work, args=(ab, bb, ref_b, 12)),
          threading.Thread(target=work, args=(aa, ba, ref_a, 12))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


def test_new_batch_needs_its_own_parameters():
    """gsm_upload_trees starts a new batch: per-tree parameters of the previous one are not silently reused."""
    b = synthetic.make_batch("C2", 8, 20)
    a = host_arrays(b)
    eng = GsmEngine(max_trees=16, max_nodes=b.n_nodes)
    run_capi(a, b.flags, b.stub_mode, engine=eng)
    eng.upload_trees(np.concatenate([a["height"], a["height"]]), np.concatenate([a["parent"], a["parent"]]))
    eng.upload_branch_params(np.concatenate([a["rate"]] * 2), np.concatenate([a["spike"]] * 2), np.concatenate([a["nstubs"]] * 2))
    with pytest.raises(GsmError) as ei:
        eng.eval()
    assert ei.value.code == abi.E_STATE
    eng.close()


def test_nonpositive_node_counts_in_a_ragged_batch():
    """node_count <= 0 gives a NaN row and does not disturb the trees evaluated after it (redo list of the lean path)."""
    b = synthetic.make_batch("C3", 64, 30)
    a = host_arrays(b)
    nc = np.full(64, b.n_nodes, np.int32)
    nc[::5] = 0
    nc[3] = -2
    a["spike"][1::5, 2] = -1.0                       # invalid states next to them: handed to the generic path as well
    ref = run_oracle(a, b.flags, b.stub_mode, node_count=nc)
    got = run_capi(a, b.flags, b.stub_mode, node_count=nc)
    assert np.all(np.isnan(got["out"][nc <= 0]))
    ok = nc > 0
    assert_close(got["out"][ok], ref["out"][ok], what="rows next to empty trees")


def test_whole_tree_stub_cdf_and_choice_match_oracle():
    """Log-time stub sampling for every branch of whole trees in two launches (gsm_stub_cdf_trees) and the draw itself for
    caller-supplied uniforms (gsm_stub_choose) against BSP.getCumulativeProbs + Randomizer.randomChoice restated in the
    oracle, and against the one-branch call."""
    b = synthetic.make_batch("C3", 6, 60, seed=4)
    a = host_arrays(b)
    flags = b.flags | abi.F_BSP_HAS_INDICATOR
    a["use_spike"][:] = [1, 0, 1, 1, 0, 1]
    eng = GsmEngine(max_trees=6, max_nodes=b.n_nodes)
    run_capi(a, flags, abi.STUBS_NOSTUB, engine=eng)
    sel = np.array([4, 0, 5, 2], np.int64)
    ptr, cdf = eng.stub_cdf_trees(sel)
    N = b.n_nodes
    assert ptr.shape == (len(sel) * N + 1,) and ptr[-1] == len(cdf)
    rng = np.random.default_rng(8)
    u = rng.random((len(sel), N))
    u[0, :5] = [0.0, 1.0, 0.999999999, 1e-300, 0.5]
    picks = eng.stub_choose(u, sel)
    L = oracle.lib()
    checked = drawn = 0
    for s, t in enumerate(sel):
        lam, mu_, psi, rho = a["params"][t, abi.P_LAMBDA:abi.P_RHO + 1]
        T_root = a["height"][t].max()
        for node in range(N):
            j = s * N + node
            got = cdf[ptr[j]:ptr[j + 1]]
            p = a["parent"][t, node]
            mean = 0.0 if p < 0 else L.gsmo_mean_stub_number(a["height"][t, node], a["height"][t, p], lam, mu_, psi, rho, T_root)
            if p < 0 or mean == 0:
                assert len(got) == 0 and picks[s, node] == 0     # no draw: root / zero-length branch (Stubs:416-434)
                continue
            ref, n_ref = oracle.cumulative_probs(mean, a["spike"][t, node], a["params"][t, abi.P_SPIKE_SHAPE], a["use_spike"][t])
            assert len(got) == n_ref
            if a["spike"][t, node] == 0 and a["use_spike"][t]:
                continue                                          # 0/0 in the Java as well
            assert_close(got, ref, rtol=1e-9, atol=1e-12, what="cdf t=%d node=%d" % (t, node))
            want = oracle.random_choice(ref, u[s, node])
            near = np.any(np.abs(ref - u[s, node]) < 1e-9)       # a uniform this close to a CDF step may fall either side
            assert picks[s, node] == want or near, (t, node, picks[s, node], want)
            checked += 1
            drawn += picks[s, node] > 0
            if node % 11 == 0:
                one, n_one = eng.stub_cdf(int(t), node)
                assert n_one == n_ref and np.array_equal(one, got)
    assert checked > 150 and drawn > 20
    # stubs estimated: the reference returns -1 for every branch (Stubs:412)
    eng.set_mode(abi.STUBS_COUNTS, flags)
    assert np.all(eng.stub_choose(u, sel) == -1)
    eng.close()


def _run_with(opts, a, flags, mode, node_count=None):
    T, N = a["height"].shape
    eng = GsmEngine(max_trees=max(T, 1), max_nodes=max(N, 1))
    for k, v in opts.items():
        eng.set_option(k, v)
    got = run_capi(a, flags, mode, node_count=node_count, engine=eng)
    eng.close()
    return got


@pytest.mark.parametrize("opts", [{"producer": 1}, {"persistent": 1}], ids=["producer-warp", "persistent-kernel"])
def test_launch_variants_of_the_lean_kernel_give_the_default_rows(opts):
    """The lean kernel's launch variants (producer warp: gsm_set_option("producer", 1); persistent kernel gsm_lean_pk.cuh:
    "persistent", 1; both off by default) do the same arithmetic in the same order as the default launch: bit-identical
    rows, rates and weighted spikes on ordinary trees, sampled-ancestor trees and ragged batches; on the case table
    (invalid states, roots that are not the last node, ...) bit-identical for the producer variant and within the parity
    tolerance for the persistent kernel, which evaluates a tree whose root is not node n-1 with the GENERAL bodies where
    the default hands the root's pair to the tail lanes."""
    bitwise_cases = "persistent" not in opts
    batches = [("C4", 700, 2000), ("C3", 600, 1000), ("C2", 1500, 200), ("C3", 40, 3), ("C4", 5, 2)]
    for kind, T, n in batches:
        b = synthetic.make_batch(kind, T, n)
        a = host_arrays(b)
        ref = run_oracle(a, b.flags, b.stub_mode)
        base = _run_with({"warp_kernel": 0}, a, b.flags, b.stub_mode)        # the CTA kernel also for small trees
        got = _run_with(dict(opts, warp_kernel=0), a, b.flags, b.stub_mode)
        assert_results_match(got, ref, what="%s %dx%d %s" % (kind, T, n, opts))
        for key in ("out", "rates", "wspikes"):
            assert np.array_equal(got[key], base[key], equal_nan=True), (kind, T, n, key)
    for name, a, flags, mode, nc in CASES:
        if mode == abi.STUBS_NOSTUB and (flags & abi.F_STUBS_TO_SPIKE_PRIOR):
            continue                                      # mixture-mode wiring: its own kernel flavour, no variants
        if a.get("stubs") is not None:
            continue
        base = _run_with({"warp_kernel": 0}, a, flags, mode, node_count=nc)
        got = _run_with(dict(opts, warp_kernel=0), a, flags, mode, node_count=nc)
        if bitwise_cases:
            for key in ("out", "rates", "wspikes"):
                assert np.array_equal(got[key], base[key], equal_nan=True), (name, key)
        else:
            assert_results_match(got, base, node_count=nc, what=name + " persistent vs default")


def _move_root(a, j):
    """The same trees with node labels n-1 (the root of the synthetic batches) and j exchanged: BEAST's operators leave the
    root at any internal node number."""
    a = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    n1 = a["height"].shape[1] - 1
    for key in ("height", "rate", "spike", "nstubs", "parent"):
        x = a[key]
        t = x[:, n1].copy()
        x[:, n1] = x[:, j]
        x[:, j] = t
    p = a["parent"]
    is1, isj = p == n1, p == j
    p[is1] = j
    p[isj] = n1
    return a


@pytest.mark.parametrize("kind,T,n", [("C4", 300, 2000), ("C2", 400, 200), ("C3", 300, 500), ("C4", 64, 501), ("C2", 64, 6)])
def test_root_at_any_internal_node_matches_oracle(kind, T, n):
    """Ordinary trees whose root is not node n-1 keep the CLEAN bodies of the lean kernel: the pair that holds the root is
    handed to two tail lanes.  Root moved to the first / second internal node (the pair that straddles the leaf boundary
    when the leaf count is odd), to n-2, n-3 and to the middle; with and without the producer warp."""
    b = synthetic.make_batch(kind, T, n)
    a0 = host_arrays(b)
    L, N = n, b.n_nodes
    for j in sorted({L, L + 1, N - 2, N - 3, (L + N) // 2, (L + N) // 2 + 1} & set(range(L, N - 1))):
        a = _move_root(a0, j)
        ref = run_oracle(a, b.flags, b.stub_mode)
        assert np.isfinite(ref["out"][:, :5]).all() or kind == "C3"
        assert_results_match(run_capi(a, b.flags, b.stub_mode), ref, what="%s root at %d (default launch)" % (kind, j))
        got = _run_with({"warp_kernel": 0}, a, b.flags, b.stub_mode)
        assert_results_match(got, ref, what="%s root at %d" % (kind, j))
        got_p = _run_with({"warp_kernel": 0, "producer": 1}, a, b.flags, b.stub_mode)
        for key in ("out", "rates", "wspikes"):
            assert np.array_equal(got_p[key], got[key], equal_nan=True), (kind, j, key)


def test_one_warp_per_tree_kernel_matches_oracle_and_the_cta_kernel():
    """Trees of up to 256 taxa are evaluated one warp per tree (gsm_lean_warp.cuh).  Same node arithmetic as the CTA kernel,
    another partition of the nodes over threads: against the oracle at the parity tolerance, against the CTA kernel
    (gsm_set_option("warp_kernel", 0)) at 1e-12, integer outputs and weighted spikes bit-exact; ordinary trees, sampled
    ancestors, tiny trees, odd / even leaf counts, ragged batches, roots that are not the last node, the case table."""
    on, off = {"warp_kernel": 1}, {"warp_kernel": 0}
    for kind, T, n in [("C2", 3000, 200), ("C3", 1200, 200), ("C4", 900, 256), ("C3", 500, 101), ("C2", 700, 33), ("C4", 300, 5),
                       ("C3", 100, 2), ("C2", 50, 1), ("C4", 64, 130)]:
        b = synthetic.make_batch(kind, T, n)
        a = host_arrays(b)
        ref = run_oracle(a, b.flags, b.stub_mode)
        got = _run_with(on, a, b.flags, b.stub_mode)
        assert_results_match(got, ref, what="warp kernel %s %dx%d" % (kind, T, n))
        cta = _run_with(off, a, b.flags, b.stub_mode)
        assert_close(got["out"], cta["out"], rtol=1e-12, atol=1e-11, what="warp vs CTA kernel %s %dx%d" % (kind, T, n))
        assert np.array_equal(got["wspikes"], cta["wspikes"], equal_nan=True)
        if T <= 1200:                                      # mixture-mode spike prior (the templates' wiring): its own flavour
            refm = run_oracle(a, b.flags, abi.STUBS_NOSTUB)
            gotm = _run_with(on, a, b.flags, abi.STUBS_NOSTUB)
            assert_results_match(gotm, refm, what="warp kernel, mixture mode %s %dx%d" % (kind, T, n))
            ctam = _run_with(off, a, b.flags, abi.STUBS_NOSTUB)
            assert_close(gotm["out"], ctam["out"], rtol=1e-12, atol=1e-11, what="warp vs CTA kernel, mixture mode %s %dx%d" % (kind, T, n))
        if n >= 5:
            for j in (n, n + 1, 2 * n - 3):
                if n <= j < 2 * n - 2:
                    am = _move_root(a, j)
                    assert_results_match(_run_with(on, am, b.flags, b.stub_mode), run_oracle(am, b.flags, b.stub_mode),
                                         what="warp kernel %s root at %d" % (kind, j))
    for name, a, flags, mode, nc in CASES:
        if a["height"].shape[1] > 511 or a.get("stubs") is not None:
            continue
        if name in ("c3-lambda-le-mu", "c3-lambda-eq-mu") and mode == abi.STUBS_NOSTUB:
            continue
        got = _run_with(on, a, flags, mode, node_count=nc)
        assert_results_match(got, run_oracle(a, flags, mode, node_count=nc), node_count=nc, what="warp kernel " + name)
    # ragged: 3 ... 511 nodes in one launch
    sizes = [2, 3, 26, 129, 200, 256]
    per = 64
    Nmax = 2 * max(sizes) - 1
    parts, counts = [], []
    for k, taxa in enumerate(sizes):
        b = synthetic.make_batch("C3", per, taxa, seed=synthetic.SEEDS["C3"] + k)
        a = host_arrays(b)
        pad = {}
        for key, fill in (("height", 0.0), ("parent", -1), ("rate", 1.0), ("spike", 0.0), ("nstubs", 0)):
            xx = np.full((per, Nmax), fill, dtype=a[key].dtype)
            xx[:, :b.n_nodes] = a[key]
            pad[key] = xx
        pad["params"], pad["use_spike"] = a["params"], a["use_spike"]
        parts.append(pad)
        counts.append(np.full(per, b.n_nodes, np.int32))
    a = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    nc = np.concatenate(counts)
    got = _run_with(on, a, synthetic.DEFAULT_FLAGS, abi.STUBS_COUNTS, node_count=nc)
    assert_results_match(got, run_oracle(a, synthetic.DEFAULT_FLAGS, abi.STUBS_COUNTS, node_count=nc), node_count=nc, what="warp kernel ragged")
