"""Stubs exact (reversible-jump) mode, STP:573-600 / 639-663: the oracle's restatement pinned by mpmath (50 digits) and by the
identity that ties the stub birth rate to the expected stub count; argument / state errors of the C ABI are in the GPU tier."""
import mpmath as mp
import numpy as np
import pytest

import oracle
from gammaspikemodel_b200 import abi, synthetic
from cases import build_cases
from helpers import host_arrays, run_oracle, stubs_from_counts

mp.mp.dps = 50
L = oracle.lib()


def _c1c2(lam, mu, psi, rho):
    c1 = mp.sqrt(abs((lam - mu - psi) ** 2 + 4 * lam * psi))
    c2 = -(lam - mu - 2 * lam * rho - psi) / c1
    return c1, c2


def p0_sampling(t, lam, mu, psi, rho):
    """Stadler (2010) eq. 1 = STP:506-507"""
    lam, mu, psi, rho, t = map(mp.mpf, (lam, mu, psi, rho, t))
    c1, c2 = _c1c2(lam, mu, psi, rho)
    e = mp.exp(-c1 * t)
    return (lam + mu + psi + c1 * (e * (1 - c2) - (1 + c2)) / (e * (1 - c2) + (1 + c2))) / (2 * lam)


def p0_nosampling(t, lam, mu):
    """Stadler (2010) remark 3.2 = STP:826-831"""
    lam, mu, t = map(mp.mpf, (lam, mu, t))
    e = mp.exp((lam - mu) * t)
    return mu * (e - 1) / (lam * e - mu)


PARAMS = [(1.3, 0.4, 0.25, 0.6), (2.25, 1.05, 0.45, 0.3), (0.7, 0.05, 1.5, 1.0), (3.0, 2.9, 0.01, 0.05)]


@pytest.mark.parametrize("lam,mu,psi,rho", PARAMS)
def test_stub_log_birth_rate_is_log_2_lambda_p0(lam, mu, psi, rho):
    for t in (1e-6, 0.01, 0.3, 1.0, 4.0, 17.5):
        p0 = p0_sampling(t, lam, mu, psi, rho)
        ref = mp.log(2 * mp.mpf(lam) * p0)
        got = L.gsmo_stub_log_birth_rate_sampling(t, lam, mu, psi, rho)
        # the Java sum lambda + mu + psi + c1 top/btm cancels when p0 is tiny (rho -> 1, t -> 0): its own rounding noise
        # is ~ ulp(l + m + psi) / (2 l p0) in the logarithm; the oracle follows the Java expression
        noise = 8e-16 * (lam + mu + psi) / float(2 * lam * p0)
        assert abs(got - float(ref)) <= 1e-12 * max(1.0, abs(float(ref))) + noise
        ref = mp.log(2 * mp.mpf(lam) * p0_nosampling(t, lam, mu))
        got = L.gsmo_stub_log_birth_rate_nosampling(t, lam, mu)
        # exp((l - m) t) - 1 cancels for small (l - m) t (the Java does not use expm1): noise ~ ulp(1) / ((l - m) t)
        assert abs(got - float(ref)) <= 1e-11 * max(1.0, abs(float(ref))) + 8e-16 / ((lam - mu) * t)


@pytest.mark.parametrize("lam,mu,psi,rho", PARAMS)
def test_expected_stub_count_is_the_integral_of_the_birth_rate(lam, mu, psi, rho):
    """E(branch) of STP:749-769 == integral over the branch of 2 lambda p0(t) (the point process of STP:573-600 and the
    Poisson count of STP:555-571 describe the same process): pins a16 / a17 and a18 against each other."""
    for h1, h0 in ((0.0, 0.7), (0.2, 0.21), (1.0, 6.5), (3.0, 3.0)):
        ref = mp.quad(lambda t: 2 * mp.mpf(lam) * p0_sampling(t, lam, mu, psi, rho), [h1, h0])
        got = L.gsmo_expected_stubs_sampling(h1, h0, lam, mu, psi, rho)
        assert abs(got - float(ref)) <= 1e-12 * max(1.0, abs(float(ref)))
        ref = mp.quad(lambda t: 2 * mp.mpf(lam) * p0_nosampling(t, lam, mu), [h1, h0])
        got = L.gsmo_expected_stubs_nosampling(h1, h0, lam, mu, 7.0)
        assert abs(got - float(ref)) <= 1e-10 * max(1.0, abs(float(ref)))


def _mp_exact_stub_logp(a, t, with_sampling):
    """whole-tree stubLogP of exact mode in 50-digit arithmetic (valid states only)"""
    lam, mu, psi, rho = a["params"][t, abi.P_LAMBDA:abi.P_RHO + 1]
    h, par = a["height"][t], a["parent"][t]
    ptr, br, tm = a["stubs"]
    br, tm = br[ptr[t]:ptr[t + 1]], tm[ptr[t]:ptr[t + 1]]
    total = mp.mpf(0)
    for i in range(len(h)):
        if par[i] < 0:
            continue
        h1, h0 = mp.mpf(float(h[i])), mp.mpf(float(h[par[i]]))
        f = (lambda x: p0_sampling(x, lam, mu, psi, rho)) if with_sampling else (lambda x: p0_nosampling(x, lam, mu))
        if h0 > h1:
            total -= mp.quad(lambda x: 2 * mp.mpf(lam) * f(x), [h1, h0])
        k = int((br == i).sum())
        total -= mp.log(mp.factorial(k))
        for s in np.where(br == i)[0]:
            x = h1 + mp.mpf(float(tm[s])) * (h0 - h1)
            total += mp.log(2 * mp.mpf(lam) * f(x))
    return total


@pytest.mark.parametrize("kind,with_sampling", [("C3", True), ("C2", False)])
def test_whole_tree_exact_stub_density_against_mpmath(kind, with_sampling):
    b = synthetic.make_batch(kind, 3, 12, seed=77)
    a = host_arrays(b)
    a["stubs"] = stubs_from_counts(a, 5)
    got = run_oracle(a, b.flags, abi.STUBS_EXACT)["out"]
    for t in range(3):
        ref = _mp_exact_stub_logp(a, t, with_sampling)
        assert abs(got[t, abi.O_STUB_LOGP] - float(ref)) <= 1e-10 * max(1.0, abs(float(ref))), (t, got[t], ref)


def test_exact_cases_hit_the_paths_they_name():
    cases = {c[0]: c for c in build_cases()}
    def out(name):
        _, a, flags, mode, nc = cases[name]
        return run_oracle(a, flags, mode, node_count=nc)["out"]
    assert np.all(np.isfinite(out("c3-exact")[:, abi.O_STUB_LOGP]))
    assert np.all(np.isfinite(out("c2-exact")[:, abi.O_STUB_LOGP]))
    o = out("c3-exact-stub-outside-branch")
    assert np.isneginf(o[::3, abi.O_STUB_LOGP]).sum() >= 4 and np.isfinite(o[1, abi.O_STUB_LOGP])
    o = out("c3-exact-stub-on-root")
    assert np.isneginf(o[::2, abi.O_STP_LOGP]).sum() >= 6 and np.isfinite(o[1, abi.O_STP_LOGP])
    o = out("c3-exact-stub-on-sampled-ancestor")
    assert np.isneginf(o[:, abi.O_STP_LOGP]).sum() >= 3
    assert np.all(out("c3-exact-ignore-stub-prior")[:, abi.O_STUB_LOGP] == 0)
    assert np.all(np.isneginf(out("c2-exact-mu-zero")[:, abi.O_STUB_LOGP]))
    # exact and count mode give the same spike prior / rates (same counts), different stub densities
    _, a, flags, _, _ = cases["c3-exact"]
    cnt = {k: v for k, v in cases["c3-default"][1].items()}
    oc = run_oracle(cnt, flags, abi.STUBS_COUNTS)["out"]
    ox = out("c3-exact")
    assert np.array_equal(oc[:, abi.O_SPIKE_LOGP], ox[:, abi.O_SPIKE_LOGP])
    assert not np.allclose(oc[:, abi.O_STUB_LOGP], ox[:, abi.O_STUB_LOGP])
