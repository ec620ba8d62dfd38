"""Accuracy of the table-driven fp64 log / exp the kernels use (gsm_fastmath.cuh), measured on the host build of
the same code against 80-bit long double references."""
import ctypes as C

import numpy as np

from helpers import emu


def _call(fn, x):
    x = np.ascontiguousarray(x, np.float64)
    y = np.empty_like(x)
    fn(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_int64(x.size))
    return y


def _ulps(y, ref):
    ulp = np.spacing(np.abs(ref.astype(np.float64))).astype(np.longdouble)
    return np.abs((y.astype(np.longdouble) - ref) / ulp).astype(np.float64)


def test_log_accuracy_and_special_values():
    rng = np.random.default_rng(1)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 400_000)), rng.uniform(0.5, 2, 400_000),
                        1 + rng.uniform(-1e-3, 1e-3, 200_000), 1 + rng.uniform(-1e-9, 1e-9, 50_000)])
    e = _ulps(_call(emu().gsm_emu_log_vec, x), np.log(x.astype(np.longdouble)))
    assert e.max() < 2.0 and e.mean() < 0.3
    sp = np.array([0.0, -0.0, -1.0, np.inf, np.nan, 5e-324, 1e-310, 2.2250738585072014e-308, 1.7976931348623157e308])
    with np.errstate(all="ignore"):
        assert np.array_equal(_call(emu().gsm_emu_log_vec, sp), np.log(sp), equal_nan=True)


def test_exp_accuracy_and_special_values():
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(-703, 703, 500_000), rng.uniform(-1, 1, 300_000), rng.uniform(-1e-5, 1e-5, 100_000)])
    e = _ulps(_call(emu().gsm_emu_exp_vec, x), np.exp(x.astype(np.longdouble)))
    assert e.max() < 1.5 and e.mean() < 0.4
    sp = np.array([0.0, -0.0, 709.0, 710.0, -745.0, -746.0, -708.5, np.inf, -np.inf, np.nan, 704.0, -704.0, 1e-320])
    with np.errstate(all="ignore"):
        assert np.array_equal(_call(emu().gsm_emu_exp_vec, sp), np.exp(sp), equal_nan=True)
