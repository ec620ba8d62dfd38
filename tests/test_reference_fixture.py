"""Pins of the oracle that do not come from the oracle: (1) the reference's own 36-taxon FBD tree with sampled ancestors
(examples/stubs/nrStubs.xml:34-40, committed as tests/golden/nrstubs_tree.npz by make_nrstubs_fixture.py) evaluated in
50-digit arithmetic from the published formulas the Java implements (Stadler 2010 eq. 1, 9; Stolz et al. 2024 eq. 1), for
rows a14 (psi > 0 tree density) and a16 (stub density with sampling); (2) caterpillar and balanced ultrametric trees for
rows a12 (psi = 0, rho = 1), a13 (psi = 0, rho < 1) and a17 (stub density without sampling).  The kernel-math emulator is
checked on the same inputs here; the -m gpu tier runs them through the C ABI (tests/cases.py)."""
import os

import mpmath as mp
import numpy as np
import pytest

from gammaspikemodel_b200 import abi, synthetic
from helpers import assert_close, run_emu, run_emu_lean, run_oracle

mp.mp.dps = 50
F = synthetic.DEFAULT_FLAGS
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_tree(T=1):
    z = np.load(os.path.join(HERE, "golden", "nrstubs_tree.npz"))
    h, p = z["height"], z["parent"]
    N = len(h)
    lam, r0, sp = float(z["lam"]), float(z["r0"]), float(z["sampling_proportion"])
    mu = lam / r0                                                     # STP:697-700
    psi = sp * mu / (1 - sp)                                          # STP:717-720
    rng = np.random.default_rng(36)
    a = dict(height=np.tile(h, (T, 1)), parent=np.tile(p, (T, 1)).astype(np.int32), rate=np.exp(0.2 * rng.standard_normal((T, N))),
             spike=rng.gamma(2.0, 0.3, (T, N)), nstubs=np.ones((T, N), np.int32), params=np.zeros((T, abi.NPARAM)),
             use_spike=np.ones(T, np.uint8))
    hp = h[np.maximum(p, 0)]
    zero_len = (p >= 0) & (hp == h)
    a["nstubs"][:, zero_len | (p < 0)] = 0                            # a zero-length branch must have zero stubs (STP:558-561)
    fake = np.zeros(N, bool)
    fake[p[zero_len]] = True
    sib = (p >= 0) & fake[np.maximum(p, 0)]
    a["spike"][:, zero_len] = 0.0                                     # sampled ancestors carry no spike (BSP:203-205)
    a["nstubs"][:, sib & ~zero_len] = np.maximum(a["nstubs"][:, sib & ~zero_len], 1)   # their siblings: m = n >= 1
    a["params"][:] = [1.0, 0.01, 2.0, 0.3, lam, mu, psi, 1.0, 0.0]
    return a


def mp_consts(lam, mu, psi, rho):
    lam, mu, psi, rho = map(mp.mpf, (lam, mu, psi, rho))
    c1 = mp.sqrt(abs((lam - mu - psi) ** 2 + 4 * lam * psi))          # STP:496-498
    c2 = -(lam - mu - 2 * lam * rho - psi) / c1                       # STP:500-502
    return lam, mu, psi, rho, c1, c2


def mp_tree_with_psi(h, par, lam, mu, psi, rho):
    """Stadler (2010) eq. 9 as STP:290-347 evaluates it"""
    lam, mu, psi, rho, c1, c2 = mp_consts(lam, mu, psi, rho)
    N = len(h)
    H = [mp.mpf(float(x)) for x in h]
    nchild = np.bincount(par[par >= 0], minlength=N)
    leaf = nchild == 0
    da = [bool(leaf[i] and par[i] >= 0 and h[par[i]] == h[i]) for i in range(N)]
    fake = [False] * N
    for i in range(N):
        if da[i]:
            fake[par[i]] = True
    p0 = lambda t: (lam + mu + psi + c1 * (mp.exp(-c1 * t) * (1 - c2) - (1 + c2)) / (mp.exp(-c1 * t) * (1 - c2) + (1 + c2))) / (2 * lam)
    p1 = lambda t: 4 * rho / (2 * (1 - c2 ** 2) + mp.exp(-c1 * t) * (1 - c2) ** 2 + mp.exp(c1 * t) * (1 + c2) ** 2)
    k = sum(da)
    n = sum(1 for i in range(N) if leaf[i] and not da[i] and h[i] <= 1e-10)
    m = sum(1 for i in range(N) if leaf[i] and not da[i] and h[i] > 1e-10 and (h[par[i]] - h[i]) > 1e-10)
    assert k + n + m == leaf.sum()
    T = H[int(np.where(par < 0)[0][0])]
    lp = mp.log(4) + mp.log(n) + mp.log(rho) + mp.log(lam) + (k + m) * mp.log(psi)
    lp -= mp.log(c1) + mp.log(c2 + 1) + mp.log(1 - c2 + (1 + c2) * mp.exp(c1 * T))
    for i in range(N):
        if not leaf[i] and not fake[i]:
            lp += mp.log(lam) + mp.log(p1(H[i]))
        if leaf[i] and h[i] > 1e-10 and not da[i]:
            lp += mp.log(p0(H[i])) - mp.log(p1(H[i]))
    return lp


def mp_stub_counts(h, par, ns, lam, mu, psi, rho):
    """Poisson stub counts with E of Stolz et al. eq. 1 (STP:749-754) or, for psi = 0 and rho = 1, STP:758-769"""
    lam, mu, psi, rho, c1, c2 = mp_consts(lam, mu, psi, rho)
    T = mp.mpf(float(h.max()))
    tot = mp.mpf(0)
    for i in range(len(h)):
        if par[i] < 0:
            continue
        h1, h0 = mp.mpf(float(h[i])), mp.mpf(float(h[par[i]]))
        if h0 <= h1:
            assert ns[i] == 0
            continue
        if psi == 0 and rho == 1:
            I = lambda x: 2 * (mp.log(lam * mp.exp((lam - mu) * x) - mu) + lam * (T - x))
            E = I(h1) - I(h0)
        else:
            f = ((c2 - 1) * mp.exp(-c1 * h1) - c2 - 1) / ((c2 - 1) * mp.exp(-c1 * h0) - c2 - 1)
            E = (h0 - h1) * (lam + mu + psi - c1) + 2 * mp.log(f)
        n = int(ns[i])
        tot += n * mp.log(E) - E - mp.log(mp.factorial(n))
    return tot


def test_reference_fbd_tree_psi_density_and_stub_density_against_50_digits():
    a = reference_tree(1)
    lam, mu, psi, rho = a["params"][0, abi.P_LAMBDA:abi.P_RHO + 1]
    h, par = a["height"][0], a["parent"][0]
    out = run_oracle(a, F, abi.STUBS_COUNTS)["out"][0]
    ref_tree = mp_tree_with_psi(h, par, lam, mu, psi, rho)
    ref_stub = mp_stub_counts(h, par, a["nstubs"][0], lam, mu, psi, rho)
    assert abs(out[abi.O_TREE_LOGP] - float(ref_tree)) <= 1e-11 * abs(float(ref_tree))      # a14
    assert abs(out[abi.O_STUB_LOGP] - float(ref_stub)) <= 1e-11 * abs(float(ref_stub))      # a16
    assert out[abi.O_N_EXTANT] == 20                                                        # README.md:144
    assert np.isfinite(out).all()
    for runner in (run_emu, run_emu_lean):                                                  # the kernels' algebra on the same tree
        got = runner(a, F, abi.STUBS_COUNTS)["out"][0]
        assert abs(got[abi.O_TREE_LOGP] - float(ref_tree)) <= 1e-11 * abs(float(ref_tree))
        assert abs(got[abi.O_STUB_LOGP] - float(ref_stub)) <= 1e-11 * abs(float(ref_stub))


def shaped_tree(kind, n, T_root):
    """ultrametric caterpillar / balanced tree with n leaves: (height, parent), BEAST numbering"""
    N = 2 * n - 1
    par = np.full(N, -1, np.int32)
    h = np.zeros(N)
    if kind == "caterpillar":
        for k in range(n - 1):
            node = n + k
            par[k + 1] = node
            par[k if k == 0 else n + k - 1] = node
            h[node] = T_root * (k + 1) / (n - 1)
    else:
        level = list(range(n))
        nxt, depth = n, 1
        while len(level) > 1:
            new = []
            for j in range(0, len(level) - 1, 2):
                par[level[j]] = par[level[j + 1]] = nxt
                h[nxt] = max(h[level[j]], h[level[j + 1]]) + T_root / np.ceil(np.log2(n)) * (0.8 + 0.4 * ((j // 2) % 2) * (len(level) > 2))
                new.append(nxt)
                nxt += 1
            if len(level) % 2:
                new.append(level[-1])
            level = new
        h[nxt - 1] = max(h[nxt - 1], T_root)
    return h, par


def mp_complete_rho(h, par, lam, mu):
    """STP:383-426: -sum of the branch integrals of B + sum over internal nodes of log(lambda (1 - Qt))"""
    lam, mu = mp.mpf(lam), mp.mpf(mu)
    T = mp.mpf(float(h.max()))
    B = lambda x: (T - x) * (lam - mu) - mp.log(lam * mp.exp(lam * T) - mu * mp.exp((T - x) * (lam - mu) + mu * T))
    nchild = np.bincount(par[par >= 0], minlength=len(h))
    tot = mp.mpf(0)
    for i in range(len(h)):
        hi = mp.mpf(float(h[i]))
        hp = hi if par[i] < 0 else mp.mpf(float(h[par[i]]))
        tot -= B(hi) - B(hp)
        if nchild[i]:
            e = mp.exp((lam - mu) * hi)
            qt = mu * (e - 1) / (lam * e - mu)
            tot += mp.log(lam) + mp.log(1 - qt)
    return tot


def mp_no_psi(h, par, lam, mu, rho):
    """STP:351-379"""
    lam, mu, rho = mp.mpf(lam), mp.mpf(mu), mp.mpf(rho)
    T = mp.mpf(float(h.max()))
    nchild = np.bincount(par[par >= 0], minlength=len(h))
    n = int((nchild == 0).sum())
    g = lambda x: rho * lam + (lam * (1 - rho) - mu) * mp.exp(-(lam - mu) * x)
    tot = mp.log(n) + mp.log(lam - mu) - (lam - mu) * T - mp.log(g(T))
    for i in range(len(h)):
        if nchild[i]:
            x = mp.mpf(float(h[i]))
            tot += mp.log(lam) + mp.log(rho) + 2 * mp.log(lam - mu) - (lam - mu) * x - 2 * mp.log(g(x))
    return tot


@pytest.mark.parametrize("kind,n", [("caterpillar", 9), ("balanced", 16), ("balanced", 11), ("caterpillar", 40)])
def test_ultrametric_shapes_against_50_digits(kind, n):
    h, par = shaped_tree(kind, n, 3.7)
    N = len(h)
    rng = np.random.default_rng(n)
    lam, mu = 1.4, 0.55
    a = dict(height=h[None], parent=par[None], rate=np.exp(0.1 * rng.standard_normal((1, N))), spike=rng.gamma(2.0, 0.2, (1, N)),
             nstubs=rng.poisson(0.7, (1, N)).astype(np.int32), params=np.array([[1.0, 0.01, 1.7, 0.25, lam, mu, 0.0, 1.0, 0.0]]),
             use_spike=np.ones(1, np.uint8))
    a["nstubs"][0, N - 1] = 0
    out = run_oracle(a, F, abi.STUBS_COUNTS)["out"][0]
    ref = mp_complete_rho(h, par, lam, mu)                                                  # a12
    assert abs(out[abi.O_TREE_LOGP] - float(ref)) <= 1e-10 * abs(float(ref))
    ref_stub = mp_stub_counts(h, par, a["nstubs"][0], lam, mu, 0.0, 1.0)                    # a17
    assert abs(out[abi.O_STUB_LOGP] - float(ref_stub)) <= 1e-10 * abs(float(ref_stub))
    kern = run_emu(a, F, abi.STUBS_COUNTS)["out"][0]
    assert_close(kern, out, what="kernel math on %s" % kind)
    b = {k: v.copy() for k, v in a.items()}
    b["params"][0, abi.P_RHO] = 0.35
    out = run_oracle(b, F, abi.STUBS_COUNTS)["out"][0]
    ref = mp_no_psi(h, par, lam, mu, 0.35)                                                  # a13
    assert abs(out[abi.O_TREE_LOGP] - float(ref)) <= 1e-10 * abs(float(ref))
    ref_stub = mp_stub_counts(h, par, b["nstubs"][0], lam, mu, 0.0, 0.35)                   # a16 with psi = 0
    assert abs(out[abi.O_STUB_LOGP] - float(ref_stub)) <= 1e-10 * abs(float(ref_stub))
    assert_close(run_emu(b, F, abi.STUBS_COUNTS)["out"][0], out, what="kernel math on %s, rho < 1" % kind)


def test_identities_of_the_survey_hold_on_the_shapes():
    """CompleteRho(T) + ln n == WithoutPsi(T, rho = 1)"""
    h, par = shaped_tree("balanced", 12, 2.2)
    lam, mu = 0.9, 0.2
    n = 12
    assert abs(float(mp_complete_rho(h, par, lam, mu) + mp.log(n) - mp_no_psi(h, par, lam, mu, 1.0))) < 1e-30
