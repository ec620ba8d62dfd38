"""The committed golden vectors (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py) against the
oracle and the host build of the kernel math (CPU tier), and against the C ABI on the device (-m gpu)."""
import os

import numpy as np
import pytest

from gammaspikemodel_b200 import abi
from helpers import assert_results_match, run_capi, run_emu, run_oracle

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))
KEYS = sorted({k.split("/")[0] for k in G.files if not k.startswith("junit")})


def _case(key):
    a = {k: G[key + "/" + k] for k in ("height", "parent", "rate", "spike", "nstubs", "params", "use_spike")}
    flags, mode = (int(v) for v in G[key + "/flags"])
    ref = {"out": G[key + "/out"], "rates": G[key + "/rates"], "wspikes": G[key + "/wspikes"]}
    return a, flags, mode, ref


@pytest.mark.parametrize("key", KEYS)
def test_oracle_reproduces_golden(key):
    a, flags, mode, ref = _case(key)
    got = run_oracle(a, flags, mode, threads=1)
    assert np.array_equal(got["out"], ref["out"], equal_nan=True)       # same code, same libm: bit-exact
    assert np.array_equal(got["rates"], ref["rates"])


@pytest.mark.parametrize("key", KEYS)
def test_emulated_kernel_math_matches_golden(key):
    a, flags, mode, ref = _case(key)
    assert_results_match(run_emu(a, flags, mode), ref, what=key)


def _junit(runner):
    import oracle as O
    lam, mu, psi, _ = O.param_maps(lambda_=2.25, turnover=0.46666667, sampling_proportion=0.3, rho=0.0)
    h, p = G["junit/height"], G["junit/parent"]
    vals = []
    for flags, rho in [(abi.F_ORIGIN, 0.0), (abi.F_ORIGIN | abi.F_COND_SAMPLING, 0.0), (abi.F_ORIGIN | abi.F_COND_RHO_SAMPLING, 0.5),
                       (abi.F_COND_ROOT | abi.F_COND_SAMPLING, 0.0), (abi.F_COND_ROOT | abi.F_COND_RHO_SAMPLING, 0.5)]:
        a = dict(height=h, parent=p, rate=np.ones_like(h), spike=np.ones_like(h), nstubs=np.zeros_like(p),
                 params=np.array([[1, 0.01, 2, 0.5, lam, mu, psi, rho, 10.0]]), use_spike=np.ones(1, np.uint8))
        vals.append(runner(a, flags | abi.F_IGNORE_STUB_PRIOR | abi.F_HAS_RATES, abi.STUBS_COUNTS)["out"][0, abi.O_STP_LOGP])
    return np.array(vals)


def test_junit_goldens_through_emulated_kernel_math():
    assert np.allclose(_junit(run_emu), G["junit/expected"], atol=1e-4, rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("key", KEYS)
def test_device_matches_golden(key):
    a, flags, mode, ref = _case(key)
    assert_results_match(run_capi(a, flags, mode), ref, what=key)


@pytest.mark.gpu
def test_junit_goldens_on_device():
    """StumpedTreePriorTest.java:35,53,71,96,114 through the C ABI (tolerance of the JUnit test: 1e-4)."""
    assert np.allclose(_junit(run_capi), G["junit/expected"], atol=1e-4, rtol=0)
