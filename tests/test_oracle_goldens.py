"""The CPU oracle against everything that pins it: the reference's five JUnit goldens
(test/gammaspike/distribution/StumpedTreePriorTest.java:35,53,71,96,114), algebraic identities, and
independent libraries (scipy.stats, mpmath at 50 digits) for the parts the reference leaves unpinned."""
import math

import mpmath as mp
import numpy as np
import pytest
import scipy.stats as st

import oracle as O

L = O.lib


# the JUnit tree ((3:1.5,4:0.5):1,(1:2,2:1):3); TreeParser offset=1 -> taxon k is node k-1
H = np.array([[0.0, 1.0, 2.5, 3.5, 4.0, 2.0, 5.0]])
P = np.array([[5, 5, 4, 4, 6, 6, -1]], dtype=np.int32)


def junit_value(flags, rho, origin=10.0):
    lam, mu, psi, _ = O.param_maps(lambda_=2.25, turnover=0.46666667, sampling_proportion=0.3, rho=rho)
    params = np.array([[1, 0.01, 2, 0.5, lam, mu, psi, rho, origin]])
    one = np.ones_like(H)
    r = O.eval_batch(flags | O.F_IGNORE_STUB_PRIOR, O.STUBS_COUNTS, H, P, one, one, np.zeros_like(P), params)
    return r["out"][0, O.O_STP_LOGP]


@pytest.mark.parametrize("flags,rho,expected", [
    (O.F_ORIGIN, 0.0, -28.7545),                                 # StumpedTreePriorTest.java:35
    (O.F_ORIGIN | O.F_COND_SAMPLING, 0.0, -28.3144),             # :53
    (O.F_ORIGIN | O.F_COND_RHO_SAMPLING, 0.5, -30.6927),         # :71
    (O.F_COND_ROOT | O.F_COND_SAMPLING, 0.0, -17.9467),          # :96
    (O.F_COND_ROOT | O.F_COND_RHO_SAMPLING, 0.5, -20.1364),      # :114
])
def test_junit_goldens(flags, rho, expected):
    assert junit_value(flags, rho) == pytest.approx(expected, abs=1e-4)   # the JUnit tolerance


def test_param_maps_match_junit_setup():
    lam, mu, psi, rho = O.param_maps(lambda_=2.25, turnover=0.46666667, sampling_proportion=0.3, rho=0.0)
    assert mu == pytest.approx(1.05, abs=1e-7) and psi == pytest.approx(0.45, abs=1e-7) and rho == 0.0
    lam2, mu2, psi2, rho2 = O.param_maps(net_div=1.0, turnover=0.5)      # STP:673-676, 693-696, 713-715, 733-735
    assert (lam2, mu2, psi2, rho2) == (2.0, 1.0, 0.0, 1.0)
    lam3, mu3, _, _ = O.param_maps(lambda_=3.0, r0=1.5)                  # STP:697-700
    assert mu3 == 2.0


def test_log_gamma_vs_mpmath_and_libm():
    xs = np.concatenate([np.linspace(0.2, 5, 97), np.linspace(5, 1000, 211)])
    for x in xs:
        ref = float(mp.loggamma(mp.mpf(float(x))))
        got = L().gsmo_log_gamma(float(x))
        assert got == pytest.approx(ref, rel=5e-14, abs=5e-14)
        assert got == pytest.approx(math.lgamma(x), rel=1e-12, abs=1e-12)   # libm agrees
    assert math.isnan(L().gsmo_log_gamma(0.0)) and math.isnan(L().gsmo_log_gamma(-1.0))


def test_gamma_log_density_vs_scipy():
    rng = np.random.default_rng(3)
    for _ in range(300):
        shape = rng.uniform(0.2, 8)
        m = rng.integers(1, 12)
        x = rng.gamma(shape * m, 1 / shape)
        got = L().gsmo_gamma_log_density(shape * m, 1 / shape, x)
        assert got == pytest.approx(st.gamma.logpdf(x, a=shape * m, scale=1 / shape), rel=1e-11, abs=1e-11)
    # x == 0 special values of the commons-math form
    assert L().gsmo_gamma_log_density(2.0, 0.5, 0.0) == -math.inf
    assert L().gsmo_gamma_log_density(0.5, 0.5, 0.0) == math.inf
    assert math.isnan(L().gsmo_gamma_log_density(1.0, 0.5, 0.0))
    assert L().gsmo_gamma_log_density(2.0, 0.5, -1.0) == -math.inf


def test_lognormal_log_density_vs_scipy():
    rng = np.random.default_rng(4)
    for _ in range(300):
        s = rng.uniform(0.05, 1.5)
        m = -0.5 * s * s
        x = float(np.exp(rng.normal(m, s)))
        got = L().gsmo_lognormal_log_density(m, s, x)
        assert got == pytest.approx(st.lognorm.logpdf(x, s, scale=math.exp(m)), rel=1e-11, abs=1e-11)
    assert L().gsmo_lognormal_log_density(0.0, 1.0, 0.0) == -math.inf


def test_p1_q_relation_and_expected_stub_identity():
    rng = np.random.default_rng(5)
    for _ in range(200):
        lam = rng.uniform(0.5, 3)
        mu = lam * rng.uniform(0.05, 0.95)
        psi = rng.uniform(0.01, 1)
        rho = rng.uniform(0.05, 1)
        h = rng.uniform(0, 5)
        c1, c2 = L().gsmo_c1(lam, mu, psi), L().gsmo_c2(lam, mu, psi, rho)
        assert L().gsmo_p1(h, rho, c1, c2) == pytest.approx(4 * rho / L().gsmo_q(h, c1, c2), rel=1e-13)   # STP:489 vs 523
        # identity (ii) of SURVEY §4: the two stub-expectation forms agree at psi = 0, rho = 1
        h1, h0 = sorted(rng.uniform(0, 5, 2))
        T = h0 + rng.uniform(0, 3)
        a = L().gsmo_expected_stubs_nosampling(h1, h0, lam, mu, T)
        b = L().gsmo_expected_stubs_sampling(h1, h0, lam, mu, 0.0, 1.0)
        assert a == pytest.approx(b, rel=1e-9, abs=1e-12)


def _mp_expected_stubs(h1, h0, lam, mu, psi, rho):
    mp.mp.dps = 50
    lam, mu, psi, rho, h1, h0 = (mp.mpf(float(v)) for v in (lam, mu, psi, rho, h1, h0))
    c1 = mp.sqrt(abs((lam - mu - psi) ** 2 + 4 * lam * psi))
    c2 = -(lam - mu - 2 * lam * rho - psi) / c1
    f = ((c2 - 1) * mp.exp(-c1 * h1) - c2 - 1) / ((c2 - 1) * mp.exp(-c1 * h0) - c2 - 1)
    return float((h0 - h1) * (lam + mu + psi - c1) + 2 * mp.log(f))


def test_expected_stubs_vs_mpmath():
    rng = np.random.default_rng(6)
    for _ in range(100):
        lam = rng.uniform(0.5, 3)
        mu = lam * rng.uniform(0.05, 0.95)
        psi = rng.uniform(0.0, 1)
        rho = rng.uniform(0.05, 1)
        h1 = rng.uniform(0, 5)
        h0 = h1 + rng.uniform(0.01, 3)
        ref = _mp_expected_stubs(h1, h0, lam, mu, psi, rho)
        assert L().gsmo_expected_stubs_sampling(h1, h0, lam, mu, psi, rho) == pytest.approx(ref, rel=1e-9, abs=1e-13)


def _caterpillar(n, T=3.0):
    """ultrametric caterpillar tree with n extant leaves: heights / parents arrays"""
    N = 2 * n - 1
    h = np.zeros(N)
    p = np.full(N, -1, np.int32)
    ih = np.linspace(T / (n - 1), T, n - 1)
    h[n:] = ih
    p[0] = n
    p[1] = n
    for k in range(1, n - 1):
        p[n + k - 1] = n + k
        p[k + 1] = n + k
    return h[None, :], p[None, :]


@pytest.mark.parametrize("n", [3, 4, 9])
def test_complete_rho_identity(n):
    """identity (i) of SURVEY §4: CompleteRho(T) + ln n == WithoutPsiSampling(T, rho -> 1)."""
    h, p = _caterpillar(n)
    one = np.ones_like(h)
    z = np.zeros_like(p)
    lam, mu = 1.7, 0.6
    base = [1, 0.01, 2, 0.5, lam, mu, 0.0]
    a = O.eval_batch(O.F_IGNORE_STUB_PRIOR, O.STUBS_COUNTS, h, p, one, one, z, np.array([base + [1.0, 0]]))["out"][0, O.O_TREE_LOGP]
    rho = 1 - 1e-12
    b = O.eval_batch(O.F_IGNORE_STUB_PRIOR, O.STUBS_COUNTS, h, p, one, one, z, np.array([base + [rho, 0]]))["out"][0, O.O_TREE_LOGP]
    assert a + math.log(n) == pytest.approx(b, rel=1e-9)


def test_cumulative_probs_properties():
    """BSP.getCumulativeProbs: prior branch reproduces the truncated, renormalised Poisson CDF (BSP:360-362)."""
    mu = 2.3
    cdf, n = O.cumulative_probs(mu, spike=0.7, shape=2.0, use_indicator_branch=False)
    pm = st.poisson.pmf(np.arange(n), mu)
    assert np.cumsum(pm)[n - 2] < 0.999 <= np.cumsum(pm)[n - 1]          # the term that crosses 0.999 is included
    assert np.allclose(cdf, np.cumsum(pm) / pm.sum(), rtol=1e-12)
    cdf2, n2 = O.cumulative_probs(mu, spike=0.7, shape=2.0, use_indicator_branch=True)
    w = pm[:n2] * st.gamma.pdf(0.7, a=2.0 * (np.arange(n2) + 1), scale=0.5)
    assert n2 == n and np.allclose(cdf2, np.cumsum(w) / w.sum(), rtol=1e-10)
    assert cdf2[-1] == pytest.approx(1.0, abs=1e-15)
