"""-m gpu: the sm_100a path, called through the C ABI (libgsm_b200.so), against the CPU oracle on the same
seeded inputs.  Tolerances: 1e-10 relative for fp64 log-densities / rates / ProportionOfSaltation (north_star),
bit-exact for integer outputs (nExtant) and for values that are pure selections (weighted spikes)."""
import numpy as np
import pytest

from gammaspikemodel_b200 import abi, synthetic
from helpers import assert_results_match, host_arrays, run_capi, run_oracle, variant

pytestmark = pytest.mark.gpu

F = synthetic.DEFAULT_FLAGS


@pytest.mark.parametrize("kind,T,n", [("C2", 256, 200), ("C3", 128, 200), ("C4", 64, 500), ("C3", 32, 1000), ("C4", 8, 2000),
                                      ("C2", 300, 5), ("C3", 100, 2), ("C2", 50, 1)])
def test_default_wiring_matches_oracle(kind, T, n):
    b = synthetic.make_batch(kind, T, n)
    a = host_arrays(b)
    ref = run_oracle(a, b.flags, b.stub_mode)
    got = run_capi(a, b.flags, b.stub_mode)
    assert_results_match(got, ref, what="%s %dx%d" % (kind, T, n))


def test_logp_only_path_matches_rates_path():
    b = synthetic.make_batch("C3", 64, 300)
    a = host_arrays(b)
    ref = run_oracle(a, b.flags, b.stub_mode)
    got = run_capi(a, b.flags, b.stub_mode, rates=False)
    assert got["rates"] is None
    assert_results_match(got, ref, what="logp-only")
