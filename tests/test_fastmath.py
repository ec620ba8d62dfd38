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


def test_lean_log6_exp6_accuracy():
    """The lean kernel's 64-entry-table log / exp (gsm_lean.cuh): absolute error budgets of its header comments."""
    rng = np.random.default_rng(3)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 300_000)), rng.uniform(0.4, 2.5, 400_000),
                        1 + rng.uniform(-1e-3, 1e-3, 100_000), np.array([1.0, 0.9945993423461914, 0.5, 2.0])])
    ref = np.log(x.astype(np.longdouble))
    for which, tol in ((0, 1.0e-13), (1, 1.2e-14)):
        y = np.empty_like(x)
        emu().gsm_emu_log6_vec(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_int64(x.size), C.c_int32(which))
        err = np.abs(y.astype(np.longdouble) - ref).astype(np.float64)
        # K ln2/64 is one rounded product: allow its rounding on top of the polynomial's budget
        assert np.all(err <= tol + 4e-16 * np.abs(ref.astype(np.float64))), (which, err.max())
    y1 = np.empty(1)
    one = np.array([1.0])
    emu().gsm_emu_log6_vec(one.ctypes.data_as(C.c_void_p), y1.ctypes.data_as(C.c_void_p), C.c_int64(1), C.c_int32(0))
    assert abs(y1[0]) < 1e-13
    x = np.concatenate([rng.uniform(-699, 0, 400_000), rng.uniform(-1, 1, 200_000), rng.uniform(-1e-5, 1e-5, 50_000)])
    y = _call(emu().gsm_emu_exp6_vec, x)
    ref = np.exp(x.astype(np.longdouble))
    rel = np.abs((y.astype(np.longdouble) - ref) / ref).astype(np.float64)
    assert np.all(rel <= 1.2e-14 + 1.2e-16 * np.abs(x)), rel.max()
