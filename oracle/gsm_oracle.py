"""ctypes binding of the CPU oracle (oracle/gsm_oracle.c).  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The oracle is the plain-C restatement of the reference's Java per-branch compute
(PunctuatedRelaxedClockModel / BranchSpikePrior / BranchRatePrior / StumpedTreePrior /
SaltativeProportionLogger); see gsm_oracle.h for citations and the parity-pinning statement.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgsm_oracle.so")

# flags / enums (mirror gsm_oracle.h)
F_HAS_INDICATOR = 1 << 0
F_NO_SPIKE_DATED_TIPS = 1 << 1
F_HAS_RATES = 1 << 2
F_RELAXED_FALSE = 1 << 3
F_STUBS_TO_CLOCK = 1 << 4
F_STUBS_TO_SPIKE_PRIOR = 1 << 5
F_STUBS_TO_TREE_PRIOR = 1 << 6
F_IGNORE_TREE_PRIOR = 1 << 7
F_IGNORE_STUB_PRIOR = 1 << 8
F_COND_SAMPLING = 1 << 9
F_COND_RHO_SAMPLING = 1 << 10
F_COND_ROOT = 1 << 11
F_ORIGIN = 1 << 12
F_BSP_HAS_INDICATOR = 1 << 13
STUBS_NOSTUB, STUBS_COUNTS, STUBS_EXACT = 0, 1, 2
NPARAM, NOUT = 9, 8
(P_CLOCK_RATE, P_SPIKE_MEAN, P_SPIKE_SHAPE, P_CLOCK_SD, P_LAMBDA, P_MU, P_PSI, P_RHO, P_ORIGIN) = range(9)
(O_STP_LOGP, O_TREE_LOGP, O_STUB_LOGP, O_SPIKE_LOGP, O_RATE_LOGP, O_SUM_BURST, O_SUM_DIST, O_N_EXTANT) = range(8)


class _Config(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("stub_mode", C.c_int32)]


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "gsm_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "libgsm_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        d = C.c_double
        for name, nargs in [("gsmo_log_gamma", 1), ("gsmo_gamma_log_density", 3), ("gsmo_normal_log_density", 3),
                            ("gsmo_lognormal_log_density", 3), ("gsmo_q", 3), ("gsmo_one_minus_p0hat", 4),
                            ("gsmo_c1", 3), ("gsmo_c2", 4), ("gsmo_p0", 6), ("gsmo_p1", 4),
                            ("gsmo_log_qt_nosampling", 3), ("gsmo_log_qt_sampling", 5),
                            ("gsmo_expected_stubs_sampling", 6), ("gsmo_expected_stubs_nosampling", 5),
                            ("gsmo_mean_stub_number", 7)]:
            f = getattr(L, name)
            f.restype = d
            f.argtypes = [d] * nargs
        L.gsmo_param_maps.restype = None
        L.gsmo_param_maps.argtypes = [d] * 7 + [C.POINTER(d)]
        L.gsmo_cumulative_probs.restype = C.c_int32
        L.gsmo_cumulative_probs.argtypes = [d, d, d, C.c_int, C.POINTER(d), C.c_int32]
        L.gsmo_eval_batch.restype = C.c_int
        L.gsmo_eval_batch.argtypes = [C.POINTER(_Config), C.c_int64, C.c_int32, C.c_int64] + [C.c_void_p] * 11 + [C.c_int]
        L.gsmo_eval_batch_stubs.restype = C.c_int
        L.gsmo_eval_batch_stubs.argtypes = [C.POINTER(_Config), C.c_int64, C.c_int32, C.c_int64] + [C.c_void_p] * 14 + [C.c_int]
        for name, nargs in [("gsmo_stub_log_birth_rate_sampling", 5), ("gsmo_stub_log_birth_rate_nosampling", 3)]:
            f = getattr(L, name)
            f.restype = d
            f.argtypes = [d] * nargs
        L.gsmo_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(a, dtype, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None:
        assert a.shape == shape, (a.shape, shape)
    return a


def eval_batch(flags, stub_mode, height, parent, rate, spike, nstubs, params, use_spike=None, node_count=None,
               want_rates=False, want_wspikes=False, n_threads=1, stubs=None):
    """Evaluate a [T][stride] batch.  Returns dict(out=[T][8], rates=[T][stride]|None, wspikes=...).
    stubs = (stub_ptr [T+1] int64, stub_branch [S] int32, stub_time [S] float64): the stub lists of EXACT mode."""
    height = _arr(height, np.float64)
    T, stride = height.shape
    parent = _arr(parent, np.int32, (T, stride))
    rate = _arr(rate, np.float64, (T, stride))
    spike = _arr(spike, np.float64, (T, stride))
    nstubs = _arr(nstubs, np.int32, (T, stride))
    params = _arr(params, np.float64, (T, NPARAM))
    use_spike = _arr(use_spike, np.uint8, (T,))
    node_count = _arr(node_count, np.int32, (T,))
    out = np.empty((T, NOUT), np.float64)
    rates = np.full((T, stride), np.nan) if want_rates else None
    wsp = np.full((T, stride), np.nan) if want_wspikes else None
    cfg = _Config(int(flags), int(stub_mode))
    sp = sb = st = None
    if stubs is not None:
        sp = _arr(stubs[0], np.int64, (T + 1,))
        sb = _arr(stubs[1], np.int32, (int(sp[-1]),))
        st = _arr(stubs[2], np.float64, (int(sp[-1]),))
    rc = lib().gsmo_eval_batch_stubs(C.byref(cfg), T, stride, stride, _ptr(height), _ptr(parent), _ptr(rate), _ptr(spike),
                                     _ptr(nstubs), _ptr(params), _ptr(use_spike), _ptr(node_count), _ptr(sp), _ptr(sb),
                                     _ptr(st), _ptr(out), _ptr(rates), _ptr(wsp), int(n_threads))
    if rc != 0:
        raise RuntimeError("gsmo_eval_batch failed: %d" % rc)
    return {"out": out, "rates": rates, "wspikes": wsp}


def cumulative_probs(mu, spike, shape, use_indicator_branch, cap=4096):
    buf = (C.c_double * cap)()
    n = lib().gsmo_cumulative_probs(mu, spike, shape, int(bool(use_indicator_branch)), buf, cap)
    return np.array(buf[: min(n, cap)], dtype=np.float64), n


def random_choice(cf, u):
    cf = np.ascontiguousarray(cf, np.float64)
    L = lib()
    L.gsmo_random_choice.restype = C.c_int32
    L.gsmo_random_choice.argtypes = [C.c_void_p, C.c_int32, C.c_double]
    return int(L.gsmo_random_choice(_ptr(cf), len(cf), float(u)))


def param_maps(lambda_=None, net_div=None, r0=None, turnover=None, psi=None, sampling_proportion=None, rho=None):
    nan = float("nan")
    v = [nan if x is None else float(x) for x in (lambda_, net_div, r0, turnover, psi, sampling_proportion, rho)]
    out = (C.c_double * 4)()
    lib().gsmo_param_maps(*v, out)
    return tuple(out)


def max_threads():
    return lib().gsmo_max_threads()
