"""Hand-built micro-trees that exercise every Node predicate combination and the quirks listed in
SURVEY.md §8a, against values worked out by hand from the Java source (cited inline)."""
import math

import numpy as np
import pytest
import scipy.stats as st

import oracle as O

F_ALL = O.F_HAS_INDICATOR | O.F_HAS_RATES | O.F_STUBS_TO_CLOCK | O.F_STUBS_TO_SPIKE_PRIOR | O.F_STUBS_TO_TREE_PRIOR

# 4 leaves: 0 extant, 1 dated, 2 sampled ancestor (height == parent 5), 3 sibling of the SA; internals 4 (parent of 0,1),
# 5 (fake: parent of 2,3), root 6.
H = np.array([[0.0, 0.4, 1.5, 0.0, 1.0, 1.5, 3.0]])
P = np.array([[4, 4, 5, 5, 6, 6, -1]], dtype=np.int32)
RATE = np.array([[1.1, 0.9, 1.3, 0.7, 1.2, 0.8, 1.0]])
SPIKE = np.array([[0.5, 1.5, 0.0, 0.0, 2.0, 0.3, 0.8]])
NST = np.array([[1, 0, 0, 0, 2, 0, 9]], dtype=np.int32)       # entry 6 (>= N-1) is ignored (Stubs:354)
PAR = np.array([[2.0, 0.01, 2.0, 0.5, 1.5, 0.5, 0.3, 0.8, 10.0]])   # clockRate, spikeMean, shape, sd, l, m, psi, rho, origin


def run(flags=F_ALL, mode=O.STUBS_COUNTS, use=1, spike=SPIKE, nst=NST, par=PAR):
    return O.eval_batch(flags, mode, H, P, RATE, spike, nst, par, np.array([use], np.uint8), want_rates=True, want_wspikes=True)


def test_rate_for_branch_and_burst():
    r = run()
    base, mean = 2.0, 0.01
    rates, ws = r["rates"][0], r["wspikes"][0]
    # PRCM:229: root, direct ancestor (zero length) -> base
    assert rates[6] == base and rates[2] == base
    # PRCM:233-236 for ordinary branches
    for i, p in [(0, 4), (1, 4), (4, 6), (5, 6)]:
        ln = H[0, p] - H[0, i]
        assert rates[i] == pytest.approx((ln * base * RATE[0, i] + SPIKE[0, i] * mean) / ln, rel=1e-15)
    # sibling of the SA (node 3): stubs wired and nstubs == 0 and parent fake -> burst 0 (PRCM:195-200)
    assert ws[3] == 0.0 and rates[3] == pytest.approx(base * RATE[0, 3], rel=1e-15)
    # weighted spikes for every node incl. root (SpikeSize:48-56)
    assert ws[6] == pytest.approx(0.8 * mean) and ws[0] == pytest.approx(0.5 * mean)
    # ProportionOfSaltation sums over len > 0, non-SA, non-root (SPL:41)
    elig = [0, 1, 3, 4, 5]
    assert r["out"][0, O.O_SUM_BURST] == pytest.approx(sum(ws[i] for i in elig), rel=1e-15)
    assert r["out"][0, O.O_SUM_DIST] == pytest.approx(sum((H[0, P[0, i]] - H[0, i]) * rates[i] for i in elig), rel=1e-15)


def test_indicator_off_and_nostub_quirk():
    assert np.all(run(use=0)["wspikes"] == 0.0)                                   # PRCM:181-183
    # nostub mode: getNStubsOnBranch == -1, so the SA-sibling burst is NOT zeroed
    sp = SPIKE.copy()
    sp[0, 3] = 0.25
    assert run(mode=O.STUBS_NOSTUB, spike=sp)["wspikes"][0, 3] == pytest.approx(0.25 * 0.01)
    # stubs not wired at all: parent fake -> 0 (PRCM:192)
    assert run(flags=O.F_HAS_INDICATOR | O.F_HAS_RATES, spike=sp)["wspikes"][0, 3] == 0.0


def test_spike_prior_known_n():
    shape = 2.0
    # m: node0 n+1=2, node1 1, node2 (SA) 0, node3 (parent fake) n=0 -> delta, node4 3, node5 1, root 0+1=1 (BSP:197-213; root count ignored)
    m = [2, 1, 0, 0, 3, 1, 1]
    exp = sum(st.gamma.logpdf(SPIKE[0, i], a=shape * m[i], scale=1 / shape) for i in range(7) if m[i] > 0)
    assert run()["out"][0, O.O_SPIKE_LOGP] == pytest.approx(exp, rel=1e-12)
    sp = SPIKE.copy()
    sp[0, 2] = 0.1                                                                # spike on an SA leaf: delta violated (BSP:86-95)
    assert run(spike=sp)["out"][0, O.O_SPIKE_LOGP] == -math.inf


def test_rate_prior():
    s = 0.5
    exp = sum(st.lognorm.logpdf(r, s, scale=math.exp(-0.5 * s * s)) for r in RATE[0])    # all N entries incl. root (BRP:43)
    assert run()["out"][0, O.O_RATE_LOGP] == pytest.approx(exp, rel=1e-12)


def test_stub_prior_counts():
    lam, mu, psi, rho = PAR[0, 4:8]
    exp = 0.0
    for i in range(6):                                                            # root excluded (STP:556)
        h1, h0 = H[0, i], H[0, P[0, i]]
        n = NST[0, i]
        if h0 <= h1:
            assert n == 0
            continue
        E = O.lib().gsmo_expected_stubs_sampling(h1, h0, lam, mu, psi, rho)
        exp += st.poisson.logpmf(n, E)
    assert run()["out"][0, O.O_STUB_LOGP] == pytest.approx(exp, rel=1e-11)
    nst = NST.copy()
    nst[0, 2] = 1                                                                 # stubs on the SA leaf (STP:232-240)
    r = run(nst=nst)["out"][0]
    assert r[O.O_STP_LOGP] == -math.inf and r[O.O_STUB_LOGP] == -math.inf


def test_tree_prior_with_psi_census():
    r = run()["out"][0]
    assert r[O.O_N_EXTANT] == 2.0                                                 # leaves 0 and 3 (STP:190-195)
    lam, mu, psi, rho = PAR[0, 4:8]
    c1, c2 = O.lib().gsmo_c1(lam, mu, psi), O.lib().gsmo_c2(lam, mu, psi, rho)
    n, m, k = 2, 1, 1                                                             # extant, extinct (node 1), SA (node 2)
    t = math.log(4) + math.log(n) + math.log(rho) + math.log(lam) + math.log(psi) * (k + m)
    t -= math.log(c1) + math.log(c2 + 1) + math.log(1 - c2 + (1 + c2) * math.exp(c1 * 3.0))
    for h in (1.0, 3.0):                                                          # internal, not fake (node 5 is fake) STP:329-334
        t += math.log(lam) + math.log(O.lib().gsmo_p1(h, rho, c1, c2))
    t += math.log(O.lib().gsmo_p0(0.4, lam, mu, psi, c1, c2)) - math.log(O.lib().gsmo_p1(0.4, rho, c1, c2))   # STP:337-343
    assert r[O.O_TREE_LOGP] == pytest.approx(t, rel=1e-13)
    assert r[O.O_STP_LOGP] == pytest.approx(r[O.O_TREE_LOGP] + r[O.O_STUB_LOGP], rel=1e-15)


def test_mixture_mode_by_hand():
    """BSP:102-172 for one branch against a direct evaluation of the truncated Poisson x Gamma sum."""
    r = run(mode=O.STUBS_NOSTUB)["out"][0, O.O_SPIKE_LOGP]
    lam, mu, psi, rho = PAR[0, 4:8]
    shape = 2.0
    tot = 0.0
    for i in range(7):
        x = SPIKE[0, i]
        root = P[0, i] < 0
        h0 = H[0, i]
        h1 = h0 if root else H[0, P[0, i]]
        mean = 0.0 if h1 <= h0 else O.lib().gsmo_expected_stubs_sampling(h0, h1, lam, mu, psi, rho)
        da = (i == 2)
        pf = (i in (2, 3))
        nsp = lambda k: k + 1 if root else (0 if da else (k if pf else k + 1))
        if mean > 0:
            cum, bp, k = 0.0, 0.0, 0
            while cum < 0.999:
                p = st.poisson.pmf(k, mean)
                cum += p
                mm = nsp(k)
                if mm == 0:
                    bp += p if x == 0 else 0.0
                elif x != 0:
                    bp += p * st.gamma.pdf(x, a=shape * mm, scale=1 / shape)
                k += 1
            tot += math.log(bp)
        else:
            mm = nsp(0)
            tot += (0.0 if x == 0 else -math.inf) if mm == 0 else st.gamma.logpdf(x, a=shape, scale=1 / shape)
    assert r == pytest.approx(tot, rel=1e-10)
