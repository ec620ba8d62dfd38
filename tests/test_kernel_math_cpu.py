"""CPU tier: the product's per-node device math (gsm_node_math.cuh / gsm_fastmath.cuh, compiled for the host
by tests/emu/gsm_emu.cpp and walked through the kernel's phases sequentially) against the oracle.  This checks
the algebraic rewrites, the flag derivation and the epilogue logic of the kernels without a GPU; the -m gpu tier
repeats the same cases through the C ABI on the device."""
import numpy as np
import pytest

from gammaspikemodel_b200 import abi
from cases import build_cases
from helpers import assert_results_match, run_emu, run_emu_fast, run_oracle

CASES = build_cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_emulated_kernel_math_matches_oracle(case):
    name, a, flags, mode, nc = case
    ref = run_oracle(a, flags, mode, node_count=nc)
    got = run_emu(a, flags, mode, node_count=nc)
    if name in ("c3-lambda-le-mu", "c3-lambda-eq-mu") and mode == abi.STUBS_NOSTUB:
        pytest.skip("mixture-mode spike prior is not defined for lambda <= mu (state is rejected by STP anyway)")
    assert_results_match(got, ref, node_count=nc, what=name)


def test_logp_only_variant_equals_rates_variant():
    name, a, flags, mode, nc = CASES[1]
    with_rates = run_emu(a, flags, mode, rates=True)
    without = run_emu(a, flags, mode, rates=False)
    # the logP-only kernel skips the division of PRCM:236; SPL's sum differs by <= 1 ulp per branch
    assert np.allclose(with_rates["out"], without["out"], rtol=1e-13, atol=0, equal_nan=True)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_emulated_fast_path_matches_oracle(case):
    """the register-hoisted fast path (gsm_eval_fast_kernel's math) on every case of its wiring family"""
    name, a, flags, mode, nc = case
    got = run_emu_fast(a, flags, mode, node_count=nc)
    if got is None:
        pytest.skip("wiring outside the fast-path family (generic kernel handles it)")
    assert_results_match(got, run_oracle(a, flags, mode, node_count=nc), node_count=nc, what=name)
