"""the tree-log writer (StumpedTreeLogger.toNewick STL:77-199 with Java Double.toString formatting) — host
code of the library, checked against a Python restatement of the Java method and against known Double.toString outputs."""
import ctypes as C

import numpy as np
import pytest

from gammaspikemodel_b200 import _lib, synthetic
from helpers import host_arrays

# Double.toString values documented in the JDK (java.lang.Double#toString) and classics of the shortest-repr literature
JAVA_DOUBLES = [(1.0, "1.0"), (100.0, "100.0"), (0.001, "0.001"), (1.0e-4, "1.0E-4"), (1.0e7, "1.0E7"), (9999999.0, "9999999.0"),
                (123456789.0, "1.23456789E8"), (0.1 + 0.2, "0.30000000000000004"), (4.9e-324, "4.9E-324"),
                (1.7976931348623157e308, "1.7976931348623157E308"), (-0.0, "-0.0"), (0.0, "0.0"), (float("nan"), "NaN"),
                (float("inf"), "Infinity"), (float("-inf"), "-Infinity"), (2.0e-3, "0.002"), (1.0e23, "1.0E23"), (5e-324 * 2 ** 10, "5.06E-321"),
                (3.0, "3.0"), (-12.5, "-12.5"), (0.5796144368376149, "0.5796144368376149"), (1234567.125, "1234567.125")]


def jd(lib, x):
    buf = C.create_string_buffer(64)
    n = lib.gsm_java_double_to_string(x, buf, 64)
    assert n == len(buf.value)
    return buf.value.decode()


def py_java_double(x):
    """independent restatement: repr() gives the shortest round-trip digits, Java's layout rules on top"""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "-0.0" if str(x).startswith("-") else "0.0"
    m, e = ("%r" % abs(x)).lower(), 0
    if "e" in m:
        m, es = m.split("e")
        e = int(es)
    ip, _, fp = m.partition(".")
    digits = (ip + fp).lstrip("0")
    e10 = e + len(ip) - 1 if ip.strip("0") else e - (len(fp) - len(fp.lstrip("0"))) - 1
    digits = digits.rstrip("0") or "0"
    if len(digits) == 1:                                             # Java: the closest two-digit decimal (4.9E-324)
        m2, e2 = ("%.1e" % abs(x)).split("e")
        if int(e2) == e10:
            digits = m2.replace(".", "").rstrip("0") or "0"
    sign = "-" if x < 0 else ""
    if 1e-3 <= abs(x) < 1e7:
        if e10 >= 0:
            d = digits.ljust(e10 + 1, "0")
            return sign + d[: e10 + 1] + "." + (d[e10 + 1:] or "0")
        return sign + "0." + "0" * (-e10 - 1) + digits
    return sign + digits[0] + "." + (digits[1:] or "0") + "E" + str(e10)


def test_java_double_to_string():
    lib = _lib.load()
    for x, want in JAVA_DOUBLES:
        assert jd(lib, x) == want, (x, jd(lib, x), want)
        assert py_java_double(x) == want
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.standard_normal(2000) * 10.0 ** rng.integers(-12, 12, 2000), rng.random(500), 1.0 / rng.integers(1, 999, 500)])
    for x in xs:
        s = jd(lib, float(x))
        assert s == py_java_double(float(x)), (x, s)
        assert float(s.replace("E", "e")) == x                       # round trip


def py_newick(a, t, left, right, nstubs, rate, metas):
    """STL:77-199 restated in Python"""
    h, par = a["height"][t], a["parent"][t]

    def rec(node):
        first = True
        if left[node] >= 0:
            s = "(" + rec(left[node])
            if right[node] >= 0:
                s += "," + rec(right[node])
            s += ")"
        else:
            s = str(node + 1)
        b2 = "[&"
        if nstubs is not None and par[node] >= 0:
            b2 += "nstubs=%d" % nstubs[node]
            first = False
        if rate is not None:
            if not first:
                b2 += ","
            b2 += "rate=" + py_java_double(float(rate[node]))
            first = False
        if metas and not first:
            b2 += ","
        for k, (mid, vals, dim) in enumerate(metas):
            if dim > node:
                b2 += mid + "=" + py_java_double(float(vals[node]))
            if len(b2) > 2 and k < len(metas) - 1:
                b2 += ","
        b2 += "]"
        if len(b2) > 3:
            s += b2
        length = 0.0 if par[node] < 0 else h[par[node]] - h[node]
        return s + ":" + py_java_double(float(length))
    return rec(int(np.where(par < 0)[0][0]))


@pytest.mark.parametrize("kind,n", [("C3", 25), ("C2", 6), ("C4", 2)])
def test_newick_writer_matches_the_restated_java_method(kind, n):
    lib = _lib.load()
    b = synthetic.make_batch(kind, 3, n, seed=5)
    a = host_arrays(b)
    N = b.n_nodes
    p = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)
    for t in range(3):
        par = a["parent"][t]
        left, right = np.full(N, -1, np.int32), np.full(N, -1, np.int32)
        for i in range(N):
            if par[i] >= 0:
                if left[par[i]] < 0:
                    left[par[i]] = i
                else:
                    right[par[i]] = i
        rate = np.ascontiguousarray(a["rate"][t] * 1.7)
        nst = np.ascontiguousarray(a["nstubs"][t])
        wsp = np.ascontiguousarray(a["spike"][t] * 0.01)
        raw = np.ascontiguousarray(a["rate"][t])
        for metas, use_stubs, use_rate in [([("GSMweightedSpikes", wsp, N), ("GSMbranchRates", raw, N - 1)], True, True),
                                           ([("GSMbranchRates", raw, N)], False, True), ([], True, False), ([], False, False),
                                           ([("x", wsp, 3), ("y", raw, N)], False, False)]:
            ids = "\n".join(m[0] for m in metas).encode()
            vals = np.ascontiguousarray(np.stack([m[1] for m in metas])) if metas else None
            dims = np.array([m[2] for m in metas], np.int32) if metas else None
            h = np.ascontiguousarray(a["height"][t])
            parc = np.ascontiguousarray(par)
            args = (N, p(left), p(right), p(parc), p(h), p(nst) if use_stubs else None, p(rate) if use_rate else None, len(metas),
                    ids if metas else None, p(vals), p(dims))
            need = lib.gsm_format_newick(*args, None, 0)
            buf = C.create_string_buffer(need + 1)
            assert lib.gsm_format_newick(*args, buf, need + 1) == need
            want = py_newick(a, t, left, right, nst if use_stubs else None, rate if use_rate else None, metas)
            assert buf.value.decode() == want
