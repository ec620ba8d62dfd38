"""Where the reference's own double-precision evaluation is ill-conditioned, 50-digit arithmetic decides who is right.

The Java computes the expected stub count of a branch as a difference of O(lambda T) numbers (STP:758-769) or as
(h0-h1)(l+m+psi-c1) + 2 log f (STP:749-754); the kernels compute it as len kE + 2 (aux - aux_parent) with aux = log G(h)
(DESIGN.md §4).  For a short branch E is tiny, the Java form loses up to ulp(lambda T) / E relative accuracy, and n log E of
a branch that carries stubs inherits it.  The kernel form must be at least as close to the exact value as the oracle
(which keeps the Java's operation order)."""
import mpmath as mp
import numpy as np

from gammaspikemodel_b200 import abi, synthetic
from helpers import host_arrays, run_emu, run_oracle, stub_term_noise

mp.mp.dps = 60


def _exact_stub_logp_nosampling(a, t):
    h, par, ns = a["height"][t], a["parent"][t], a["nstubs"][t]
    lam, mu = (mp.mpf(float(x)) for x in a["params"][t, abi.P_LAMBDA:abi.P_MU + 1])
    T = mp.mpf(float(h.max()))
    I = lambda x: 2 * (mp.log(lam * mp.exp((lam - mu) * x) - mu) + lam * (T - x))
    tot = mp.mpf(0)
    for i in range(len(h)):
        if par[i] < 0:
            continue
        h1, h0 = mp.mpf(float(h[i])), mp.mpf(float(h[par[i]]))
        if h0 <= h1:
            continue
        E = I(h1) - I(h0)
        n = int(ns[i])
        tot += n * mp.log(E) - E - mp.log(mp.factorial(n))
    return tot


def test_short_branches_with_stubs_kernel_form_is_closer_to_exact_than_the_java_order():
    b = synthetic.make_batch("C2", 6, 60, seed=31)
    a = host_arrays(b)
    T, N = a["height"].shape
    L = (N + 1) // 2
    rng = np.random.default_rng(1)
    for t in range(T):                                   # leaves hanging 1e-4 .. 1e-6 (relative) below their parents, with stubs
        for i in rng.choice(L, 8, replace=False):
            p = a["parent"][t, i]
            a["height"][t, i] = a["height"][t, p] * (1 - 10.0 ** -rng.integers(4, 7))
            a["nstubs"][t, i] = int(rng.integers(1, 4))
    ora = run_oracle(a, b.flags, b.stub_mode, rates=False)["out"][:, abi.O_STUB_LOGP]
    ker = run_emu(a, b.flags, b.stub_mode, rates=False)["out"][:, abi.O_STUB_LOGP]
    noise = stub_term_noise(a)
    assert noise.max() > 1e-9                            # the case is really ill-conditioned for the Java order
    for t in range(T):
        exact = _exact_stub_logp_nosampling(a, t)
        e_ora, e_ker = abs(mp.mpf(float(ora[t])) - exact), abs(mp.mpf(float(ker[t])) - exact)
        assert e_ora <= noise[t] + 1e-10 * abs(exact)    # the bound describes the oracle's error
        assert e_ker <= e_ora + 1e-11 * abs(exact)       # and the kernel's algebra is at least as accurate
        assert e_ker <= 0.05 * noise[t] + 1e-10 * abs(exact)   # ... by a wide margin where it matters
