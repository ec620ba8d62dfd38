"""-m gpu: BASELINE config 5 — batches of MCMC proposals on resident trees (gsm_eval_dirty).  The reference
re-evaluates every branch after every proposal (BSP:294-295, STP:842), so the expected value of a proposal is the
oracle on the proposed state.  Small proposals are evaluated incrementally (O(dirty), regrouped sums): they agree with
gsm_eval on the proposed state to the 1e-10 tolerance of the path, not bit for bit; proposals that take the full path
(tree scalers, hyper-parameter moves) are bit-identical to gsm_eval."""
import numpy as np
import pytest

from gammaspikemodel_b200 import GsmEngine, GsmError, abi, synthetic
from helpers import apply_proposals, assert_close, host_arrays, run_capi, run_oracle, stub_term_noise

pytestmark = pytest.mark.gpu


def _engine_with(a, b):
    eng = GsmEngine(max_trees=a["height"].shape[0], max_nodes=b.n_nodes)
    run_capi(a, b.flags, b.stub_mode, engine=eng, rates=False)
    return eng


@pytest.mark.parametrize("kind,T,n,P,hyper", [("C2", 8, 120, 300, True), ("C3", 8, 60, 300, True), ("C4", 4, 500, 200, False)])
def test_proposals_match_oracle_and_full_reevaluation(kind, T, n, P, hyper):
    b = synthetic.make_batch(kind, T, n)
    a = host_arrays(b)
    pr = synthetic.make_proposals(a, P, hyper=hyper)
    eng = _engine_with(a, b)
    got = eng.eval_dirty(**pr)
    eng.close()
    prop = apply_proposals(a, pr)
    assert_close(got, run_oracle(prop, b.flags, b.stub_mode, rates=False)["out"], what="%s proposals vs oracle" % kind)
    full = run_capi(prop, b.flags, b.stub_mode, rates=False)["out"]
    assert_close(got, full, what="%s proposals vs full re-evaluation" % kind)
    big = pr["height_scale"] != 1.0                      # tree scalers: materialised and evaluated by the same kernels
    assert big.any() and np.array_equal(got[big], full[big], equal_nan=True)


def test_proposals_chunked_and_resident_batch_untouched():
    b = synthetic.make_batch("C3", 16, 100)
    a = host_arrays(b)
    pr = synthetic.make_proposals(a, 500)
    eng = _engine_with(a, b)
    before = eng.eval()["out"]
    one = eng.eval_dirty(**pr)
    eng2 = _engine_with(a, b)
    eng2.set_option("host_chunk_mb", 1)                  # ~160 proposals per chunk
    many = eng2.eval_dirty(**pr)
    assert np.array_equal(one, many, equal_nan=True)
    assert np.array_equal(eng.eval()["out"], before, equal_nan=True)      # proposals do not modify the resident batch
    # an empty dirty set with unchanged scalars is the current state
    P = 5
    same = eng.eval_dirty(np.arange(P, dtype=np.int32), np.zeros(P + 1, np.int64), np.zeros(0, np.int32))
    assert_close(same, before[:P], what="empty dirty sets")
    eng.close()
    eng2.close()


def test_proposals_argument_errors():
    b = synthetic.make_batch("C2", 4, 20)
    a = host_arrays(b)
    eng = _engine_with(a, b)
    with pytest.raises(GsmError):                        # tree index outside the batch
        eng.eval_dirty(np.array([4], np.int32), np.array([0, 0], np.int64), np.zeros(0, np.int32))
    with pytest.raises(GsmError):                        # node outside the tree
        eng.eval_dirty(np.array([0], np.int32), np.array([0, 1], np.int64), np.array([b.n_nodes], np.int32), new_spike=np.array([1.0]))
    with pytest.raises(GsmError):                        # decreasing CSR pointer
        eng.eval_dirty(np.array([0, 1], np.int32), np.array([0, 2, 1], np.int64), np.array([0], np.int32))
    eng.close()


def test_config5_shape_64_chains_500_taxa():
    """BASELINE config 5: ONE 500-taxon tree replicated to 64 chain states (each chain with its own rates / spikes / stub
    counts / scalars) x 256 proposals each (a bounded slice of 64 x 4,096): every proposal against the oracle; device-bound
    resident batch; ~80 % of the proposals take the incremental path."""
    import torch
    b = synthetic.make_chain_states(64, 500)
    a = host_arrays(b)
    assert all(np.array_equal(a["parent"][0], a["parent"][c]) and np.array_equal(a["height"][0], a["height"][c]) for c in range(64))
    pr = synthetic.make_proposals(a, 64 * 256)
    d = b.to("cuda")
    eng = GsmEngine()
    eng.set_mode(b.stub_mode, b.flags)
    eng.bind_batch(d)
    got = eng.eval_dirty(**pr)
    n_full = eng.last_dirty_full_count
    eng.close()
    prop = apply_proposals(a, pr)
    ref = run_oracle(prop, b.flags, b.stub_mode, rates=False, threads=0)["out"]
    other = [abi.O_TREE_LOGP, abi.O_SPIKE_LOGP, abi.O_RATE_LOGP, abi.O_SUM_BURST, abi.O_SUM_DIST, abi.O_N_EXTANT]
    assert_close(got[:, other], ref[:, other], what="config 5")
    # stub density: 1e-10, plus the reference's own rounding noise where a proposal puts a stub on a branch with a tiny
    # expected count (a handful of the 16,384 proposals; see helpers.stub_term_noise)
    noise = stub_term_noise(prop)
    for col in (abi.O_STP_LOGP, abi.O_STUB_LOGP):
        err = np.abs(got[:, col] - ref[:, col])
        assert np.all(err <= 1e-9 + 1e-10 * np.abs(ref[:, col]) + noise), (col, np.argmax(err - noise))
    assert (noise > 1e-9).mean() < 0.01
    assert np.isfinite(got[:, abi.O_SPIKE_LOGP]).mean() > 0.5
    assert 0.1 < n_full / len(got) < 0.3                 # scalers + hyper-parameter moves; everything else is O(dirty)


@pytest.mark.parametrize("mode,flags_extra", [(abi.STUBS_COUNTS, 0), (abi.STUBS_NOSTUB, 0), (abi.STUBS_COUNTS, abi.F_COND_ROOT | abi.F_COND_SAMPLING),
                                              (abi.STUBS_COUNTS, abi.F_NO_SPIKE_DATED_TIPS)])
def test_incremental_path_on_sampled_ancestor_moves(mode, flags_extra):
    """Height edits that create and destroy sampled ancestors (a leaf moved onto / off its parent's height: isDirectAncestor
    and isFake change for the leaf, its parent and its sibling), edits of the root height, and edits that make the state
    invalid -- all on the incremental path, against the oracle; every wiring family (incl. the mixture-mode spike prior)."""
    b = synthetic.make_batch("C3", 12, 40, seed=11)
    a = host_arrays(b)
    flags = b.flags | flags_extra
    T, N = a["height"].shape
    L = (N + 1) // 2
    rng = np.random.default_rng(3)
    tree, ptr, nodes, nh, ns, nk, nr = [], [0], [], [], [], [], []
    def prop(t, edits):
        tree.append(t)
        for (i, h, s, k, r) in edits:
            nodes.append(i); nh.append(h); ns.append(s); nk.append(k); nr.append(r)
        ptr.append(len(nodes))
    for t in range(T):
        h, p = a["height"][t], a["parent"][t]
        keep = lambda i: (i, h[i], a["spike"][t, i], a["nstubs"][t, i], a["rate"][t, i])
        for i in rng.choice(L, 6, replace=False):       # leaf onto its parent's height: becomes a sampled ancestor
            i = int(i)
            prop(t, [(i, h[p[i]], 0.0, 0, a["rate"][t, i])])
            prop(t, [(i, h[p[i]], a["spike"][t, i], a["nstubs"][t, i], a["rate"][t, i])])   # ... with its spike / stubs kept: -inf paths
        sa = [i for i in range(L) if h[p[i]] == h[i]]
        for i in sa[:4]:                                 # sampled ancestor moved below its parent: an ordinary dated tip again
            prop(t, [(i, 0.5 * h[p[i]], 0.3, 1, 1.1)])
        root = int(np.where(p < 0)[0][0])
        prop(t, [(root, h[root] * 1.3, a["spike"][t, root], a["nstubs"][t, root], a["rate"][t, root])])   # root height
        ch = np.where(p == root)[0]
        prop(t, [(root, max(h[ch]) + 1e-3, 0.2, 0, 0.9), keep(int(ch[0]))])
        i = int(rng.integers(L, N - 1))
        prop(t, [(i, h[p[i]], a["spike"][t, i], 0, 1.0)])                      # zero-length internal branch
        prop(t, [(i, h[p[i]] * 1.5, a["spike"][t, i], 0, 1.0)])                # negative branch length
        prop(t, [(int(rng.integers(0, N - 1)), h[0], -0.5, 0, 1.0)])           # negative spike -> -inf (BSP:72-75)
        prop(t, [(int(rng.integers(0, N - 1)), h[1], 0.1, 0, -1.0)])           # negative rate (BRP:47-50)
        prop(t, [(int(rng.integers(0, N - 1)), h[2], 0.1, 200, 1.0)])          # a stub count beyond the small tables
    pr = dict(tree_of_prop=np.array(tree, np.int32), dirty_ptr=np.array(ptr, np.int64), dirty_nodes=np.array(nodes, np.int32),
              new_height=np.array(nh), new_spike=np.array(ns), new_nstubs=np.array(nk, np.int32), new_rate=np.array(nr))
    eng = GsmEngine(max_trees=T, max_nodes=N)
    run_capi(a, flags, mode, engine=eng, rates=False)
    got = eng.eval_dirty(**pr)
    assert eng.last_dirty_full_count == 0                # every one of them was evaluated incrementally
    eng.close()
    ref = run_oracle(apply_proposals(a, pr), flags, mode, rates=False)["out"]
    assert_close(got, ref, what="sampled-ancestor moves, mode %d" % mode)
    assert np.isneginf(ref[:, abi.O_SPIKE_LOGP]).any() and np.isfinite(ref[:, abi.O_STP_LOGP]).any()


def test_invalid_resident_state_takes_the_full_path():
    """a resident tree whose own state is invalid (negative spike: -inf sums) has no usable cache: its proposals are
    materialised and re-evaluated, including the one that repairs the state"""
    b = synthetic.make_batch("C2", 6, 30, seed=2)
    a = host_arrays(b)
    a["spike"][2, 5] = -1.0
    a["params"][4, abi.P_LAMBDA] = 0.5 * a["params"][4, abi.P_MU]     # lambda <= mu: STP -inf before looking at the tree
    eng = GsmEngine(max_trees=6, max_nodes=b.n_nodes)
    run_capi(a, b.flags, b.stub_mode, engine=eng, rates=False)
    tree = np.array([2, 2, 4, 0], np.int32)
    pr = dict(tree_of_prop=tree, dirty_ptr=np.array([0, 1, 2, 3, 4], np.int64), dirty_nodes=np.array([5, 6, 3, 1], np.int32),
              new_spike=np.array([0.4, 0.2, 0.3, 0.25]))
    got = eng.eval_dirty(**pr)
    assert eng.last_dirty_full_count in (2, 3)           # the two proposals on tree 2 (and tree 4 when lambda <= mu leaves non-finite sums)
    eng.close()
    ref = run_oracle(apply_proposals(a, pr), b.flags, b.stub_mode, rates=False)["out"]
    assert_close(got, ref, what="invalid resident state")
    assert np.isfinite(got[0, abi.O_SPIKE_LOGP]) and np.isneginf(got[1, abi.O_SPIKE_LOGP])
