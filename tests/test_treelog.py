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


def _write(lib, N, left, right, par, h, nst, rate, metas):
    p = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)
    ids = "\n".join(m[0] for m in metas).encode()
    vals = np.ascontiguousarray(np.stack([m[1] for m in metas])) if metas else None
    dims = np.array([m[2] for m in metas], np.int32) if metas else None
    args = (N, p(left), p(right), p(par), p(h), p(nst), p(rate), len(metas), ids if metas else None, p(vals), p(dims))
    need = lib.gsm_format_newick(*args, None, 0)
    buf = C.create_string_buffer(need + 1)
    assert lib.gsm_format_newick(*args, buf, need + 1) == need
    return buf.value


def _parse(lib, s, cap, keys):
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    par, left, right = (np.full(cap, -7, np.int32) for _ in range(3))
    length, height = np.zeros(cap), np.zeros(cap)
    vals = np.zeros((max(len(keys), 1), cap))
    n = lib.gsm_parse_newick(s, cap, p(par), p(left), p(right), p(length), p(height), len(keys), "\n".join(keys).encode(), p(vals))
    return n, par, left, right, length, height, vals


@pytest.mark.parametrize("kind,n", [("C3", 40), ("C2", 7), ("C4", 2), ("C4", 129)])
def test_newick_reader_inverts_the_writer_and_gives_the_effective_leaf_rates(kind, n):
    """gsm_parse_newick (what util/GetEffectiveRates.java:30-83 reads through BEAST's tree-log parser) on the writer's output:
    same tree (leaves keep their numbers, every parent / child relation and branch length survives the internal renumbering),
    same metadata to the last bit (Double.toString round-trips), heights with the youngest leaf at 0; eRate = rate + spike /
    length for the leaves as GetEffectiveRates.java:58-61 computes it."""
    lib = _lib.load()
    b = synthetic.make_batch(kind, 3, n, seed=11)
    a = host_arrays(b)
    N, L = b.n_nodes, n
    for t in range(3):
        par = np.ascontiguousarray(a["parent"][t])
        left, right = np.full(N, -1, np.int32), np.full(N, -1, np.int32)
        for i in range(N):
            if par[i] >= 0:
                if left[par[i]] < 0:
                    left[par[i]] = i
                else:
                    right[par[i]] = i
        h = np.ascontiguousarray(a["height"][t])
        nst = np.ascontiguousarray(a["nstubs"][t])
        rate = np.ascontiguousarray(a["rate"][t])
        spk = np.ascontiguousarray(a["spike"][t] * 0.013)
        s = _write(lib, N, left, right, par, h, nst, rate, [("spikes", spk, N), ("branchRates", rate, N)])
        got_n, p2, l2, r2, len2, h2, vals = _parse(lib, s + b";", N, ["nstubs", "rate", "spikes", "branchRates", "absent"])
        assert got_n == N
        # leaves keep their numbers; map the internal nodes through the parent relation of matching children
        m = {i: i for i in range(L)}
        for i in sorted(range(N), key=lambda i: h[i]):                 # children before parents
            if par[i] >= 0:
                assert m.setdefault(int(par[i]), int(p2[m[i]])) == int(p2[m[i]])
        assert len(set(m.values())) == N and p2[m[int(np.where(par < 0)[0][0])]] == -1 and m[int(np.where(par < 0)[0][0])] == N - 1
        for i in range(N):
            j = m[i]
            want_len = 0.0 if par[i] < 0 else h[par[i]] - h[i]
            assert len2[j] == want_len                                 # Double.toString round-trips exactly
            assert vals[1][j] == rate[i] and vals[2][j] == spk[i] and vals[3][j] == rate[i] and np.isnan(vals[4][j])
            assert (np.isnan(vals[0][j]) if par[i] < 0 else vals[0][j] == nst[i])
            assert {int(l2[j]), int(r2[j])} == ({-1} if left[i] < 0 else {m[int(left[i])], m[int(right[i])]})
        assert np.allclose(h2[[m[i] for i in range(N)]] - h2[:L].min(), h - h[:L].min(), rtol=0, atol=1e-9 * max(1.0, h.max()))
        out = np.zeros(L)
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        assert lib.gsm_effective_leaf_rates(L, p(vals[3]), p(vals[2]), p(len2), p(out)) == 0
        with np.errstate(divide="ignore", invalid="ignore"):
            want = rate[:L] + spk[:L] / np.array([h[par[i]] - h[i] for i in range(L)])
        assert np.array_equal(out, want, equal_nan=True)


def test_newick_reader_rejects_what_it_cannot_represent():
    lib = _lib.load()
    assert _parse(lib, b"((1:1.0,2:1.0):1.0,3:2.0);", 8, [])[0] == 5
    assert _parse(lib, b"((1:1.0,2:1.0):1.0,3:2.0);", 3, [])[0] == 5                 # arrays too small: the size needed
    for bad in (b"((1:1.0,2:1.0,3:1.0):1.0,4:2.0);", b"((1:1.0,2:1.0):1.0,5:2.0);", b"((1:1.0,1:1.0):1.0,3:2.0);", b"(1:1.0,2:1.0", b"(1:x,2:1.0);",
                b"((1[&a]:1.0,2:1.0):1.0,3:2.0);"):
        assert _parse(lib, bad, 16, ["a"])[0] == -1, bad
    n, par, left, right, length, height, vals = _parse(lib, b"((1[&r=1.0E-4,s=NaN]:0.5,2[&r=Infinity]:1.5)[&r=2.0]:1.0,3:3.0)", 8, ["r", "s"])
    assert n == 5 and list(par[:5]) == [3, 3, 4, 4, -1] and list(height[:5]) == [1.5, 0.5, 0.0, 2.0, 3.0]
    assert vals[0][0] == 1e-4 and np.isnan(vals[1][0]) and vals[0][1] == np.inf and vals[0][3] == 2.0 and np.isnan(vals[0][2])
