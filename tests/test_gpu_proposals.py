"""-m gpu: BASELINE config 5 — batches of MCMC proposals on resident trees (gsm_eval_dirty).  The reference
re-evaluates every branch after every proposal (BSP:294-295, STP:842), so the expected value of a proposal is the
oracle on the proposed state; and gsm_eval_dirty must be bit-identical to gsm_eval on that state."""
import numpy as np
import pytest

from gammaspikemodel_b200 import GsmEngine, GsmError, abi, synthetic
from helpers import apply_proposals, assert_close, host_arrays, run_capi, run_oracle

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
    assert np.array_equal(got, full, equal_nan=True)


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
    assert np.array_equal(same, before[:P], equal_nan=True)
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
    """64 chain states x 256 proposals each on 500-taxon trees (a bounded slice of 64 x 4,096): every proposal against the
    oracle; device-bound resident batch."""
    import torch
    b = synthetic.make_batch("C2", 64, 500, seed=synthetic.SEEDS["C5"])
    a = host_arrays(b)
    pr = synthetic.make_proposals(a, 64 * 256)
    d = b.to("cuda")
    eng = GsmEngine()
    eng.set_mode(b.stub_mode, b.flags)
    eng.bind_batch(d)
    got = eng.eval_dirty(**pr)
    eng.close()
    ref = run_oracle(apply_proposals(a, pr), b.flags, b.stub_mode, rates=False, threads=0)["out"]
    assert_close(got, ref, what="config 5")
    assert np.isfinite(got[:, abi.O_SPIKE_LOGP]).mean() > 0.5
