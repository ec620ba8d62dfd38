"""CPU tier: the product's per-node device math (gsm_node_math.cuh / gsm_fastmath.cuh, compiled for the host
by tests/emu/gsm_emu.cpp and walked through the kernel's phases sequentially) against the oracle.  This checks
the algebraic rewrites, the flag derivation and the epilogue logic of the kernels without a GPU; the -m gpu tier
repeats the same cases through the C ABI on the device."""
import numpy as np
import pytest

from gammaspikemodel_b200 import abi
from cases import build_cases
from helpers import assert_results_match, run_emu, run_emu_lean, run_oracle

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
def test_emulated_lean_path_matches_oracle(case):
    """the lean path (gsm_lean_kernel's math + redo hand-back) on every case of its wiring family"""
    name, a, flags, mode, nc = case
    got = run_emu_lean(a, flags, mode, node_count=nc)
    if got is None:
        pytest.skip("wiring outside the lean family (generic kernel handles it)")
    assert_results_match(got, run_oracle(a, flags, mode, node_count=nc), node_count=nc, what=name)


# cases whose trees must all be finished by the lean path itself (no hand-back): ordinary valid states.
# (Sampled ancestors under psi == 0 -- "c3-psi0-*-dated" -- are not in the list: the lean kernel only scans for sampled
# ancestors where the tree prior can generate them, psi > 0 or origin / root conditioning; elsewhere a zero-length branch
# is caught by its length check and the tree is handed to the generic path.)
LEAN_FULL = {"c2-default", "c3-default", "c4-default", "c2-nostub-mixture", "c3-nostub-mixture", "c4-nostub-mixture", "c3-strict-clock", "c3-no-indicator", "c3-ignore-tree-prior", "c3-ignore-stub-prior",
             "c3-origin", "c3-origin-cond-sampling", "c3-origin-cond-rho", "c3-root-cond-sampling", "c3-root-cond-rho",
             "c2-psi0-rho-lt1", "c2-psi-on-ultrametric", "ragged-node-counts"}


@pytest.mark.parametrize("case", [c for c in CASES if c[0] in LEAN_FULL], ids=[c[0] for c in CASES if c[0] in LEAN_FULL])
def test_lean_path_covers_ordinary_states(case):
    name, a, flags, mode, nc = case
    got = run_emu_lean(a, flags, mode, node_count=nc)
    assert got is not None and got["n_lean"] == a["height"].shape[0], (name, got and got["n_lean"])


@pytest.mark.parametrize("kind,T,n", [("C4", 12, 60), ("C2", 12, 33), ("C4", 6, 5), ("C3", 24, 40)])
def test_lean_path_keeps_trees_whose_root_is_not_the_last_node(kind, T, n):
    """BEAST's operators leave the root at any internal node number.  The lean path finishes such ordinary trees itself
    (the pair of nodes that holds the root goes to the tail lanes; tests/emu mirrors lean_b_pair / tail_eval) and agrees
    with the oracle, for the root at the first / second internal node, in the middle and at n-2 / n-3."""
    from gammaspikemodel_b200 import synthetic
    from helpers import host_arrays
    L, N = n, 2 * n - 1
    for j in sorted({L, L + 1, (L + N) // 2, (L + N) // 2 + 1, N - 3, N - 2} & set(range(L, N - 1))):
        b = synthetic.make_batch(kind, T, n)
        synthetic.move_root(b, j)
        a = host_arrays(b)
        got = run_emu_lean(a, b.flags, b.stub_mode)
        assert got is not None
        if kind != "C3":                                   # (C3 trees with sampled ancestors may be handed back)
            assert got["n_lean"] == T, (kind, j, got["n_lean"])
        assert_results_match(got, run_oracle(a, b.flags, b.stub_mode), what="%s root at %d" % (kind, j))
