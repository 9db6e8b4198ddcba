"""CPU check of the fp64 epilogue (icikt_common.cuh, the code the K3 kernel runs) against the
oracle: per-column tie statistics + (dis, ntie, b) of the GLOBAL problem must reproduce the
reference's results for both perspectives, including the degenerate statuses.
Tolerances: tau/tau_max 1e-12 relative (north_star), p-value 1e-9 relative (z^2-amplified,
SURVEY.md 7.3), completeness bit-exact (x87 long-double arithmetic reproduced with integers)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def epi(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("epi") / "libepilogue_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                           os.path.join(HERE, "epilogue_host.cpp"), "-o", out])
    L = ctypes.CDLL(out)
    lp = ctypes.POINTER(ctypes.c_longlong)
    L.epilogue_host.argtypes = [ctypes.c_longlong, lp, lp, ctypes.c_longlong, ctypes.c_longlong,
                                ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                ctypes.POINTER(ctypes.c_double), lp]
    L.epilogue_host.restype = ctypes.c_int
    return L


def col_stats(v):
    """n_na, n_groups, g0extra, s2o, s3o, s5o as K1 defines them (icikt_common.cuh ColStats)."""
    na = np.isnan(v)
    a = int(na.sum())
    vals = np.sort(v[~na])
    sizes = []
    if vals.size:
        vals = vals + 0.0  # -0.0 -> +0.0
        _, cnt = np.unique(vals, return_counts=True)
        sizes = list(cnt)
    g0extra = 0
    if a > 0 and vals.size and (vals[0] - 0.1 == vals[0]):
        g0extra = sizes.pop(0)
    t = np.array(sizes, dtype=np.int64)
    s2, s3, s5 = int((t * (t - 1)).sum()), int((t * (t - 1) * (t - 2)).sum()), int((t * (t - 1) * (2 * t + 5)).sum())
    return np.array([a, len(sizes) + (1 if a > 0 else 0), g0extra, s2, s3, s5], dtype=np.int64)


def group0(v):
    """rows in the lowest tie group: the missing rows, plus the minimum if min - 0.1 == min"""
    na = np.isnan(v)
    if na.any() and (~na).any():
        mn = np.nanmin(v)
        if mn - 0.1 == mn:
            return na | (v == mn)
    return na


def run(epi, x, y, persp, alt, cont):
    n = x.size
    g = O.ici_kt(x, y, "global")  # integer inputs of the epilogue come from the global problem
    b = int((np.isnan(x) & np.isnan(y)).sum())
    if g.status == 0:
        dis, ntie = g.dis, g.ntie
    else:
        dis, ntie = 0, 0
    xs, ys = col_stats(x), col_stats(y)
    g00 = int((group0(x) & group0(y)).sum())
    out4 = np.zeros(4)
    oc = np.zeros(4, dtype=np.int64)
    lp = ctypes.POINTER(ctypes.c_longlong)
    st = epi.epilogue_host(n, xs.ctypes.data_as(lp), ys.ctypes.data_as(lp), dis, ntie, b, g00,
                           O.PERSPECTIVE.get(persp, 0), O.ALTERNATIVE.get(alt, 3), int(cont),
                           out4.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), oc.ctypes.data_as(lp))
    return st, out4, oc


def close(a, b, rel):
    if np.isnan(a) and np.isnan(b):
        return True
    return a == b or abs(a - b) <= rel * abs(b)


@pytest.mark.parametrize("persp", ["global", "local"])
def test_epilogue_random(epi, persp):
    rng = np.random.default_rng(42)
    n_checked = 0
    for trial in range(600):
        n = int(rng.integers(2, 90))
        mode = trial % 5
        x = rng.normal(size=n)
        y = x * 0.5 + rng.normal(size=n)
        if mode in (1, 3):
            x = np.round(x * 1.5)
        if mode in (2, 3):
            y = np.round(y * 1.5)
        if mode == 4:
            x = np.full(n, 2.0) if trial % 2 else x
            y = np.where(rng.random(n) < 0.5, 1.0, np.nan)
        x[rng.random(n) < rng.choice([0, 0.3, 0.6, 1.0], p=[.3, .4, .25, .05])] = np.nan
        y[rng.random(n) < rng.choice([0, 0.3, 0.6])] = np.nan
        alt = ["two.sided", "less", "greater", "bogus"][trial % 4]
        cont = bool((trial // 4) % 2)
        ref = O.ici_kt(x, y, persp, alt, cont)
        g = O.ici_kt(x, y, "global")
        if g.status != 0 and ref.status == 0:
            continue  # cannot happen: local only removes rows
        st, out4, oc = run(epi, x, y, persp, alt, cont)
        assert st == ref.status, (trial, st, ref.status)
        if st == 0:
            n_checked += 1
            assert (oc[0], oc[1], oc[2], oc[3]) == (ref.xtie, ref.ytie, ref.tot, ref.n_entry)
            assert close(out4[0], ref.tau, 1e-12) and close(out4[2], ref.tau_max, 1e-12)
            assert close(out4[1], ref.pvalue, 1e-9), (out4[1], ref.pvalue)
            assert out4[3] == ref.completeness  # bit-exact
        else:
            assert np.isnan(out4).all()
    assert n_checked > 200


def test_epilogue_absorbed_minimum(epi):
    # |min| so large that min - 0.1 == min: missing rows tie with the minimum
    x = np.array([1e17, np.nan, 2e17, 1e17, np.nan, 3e17, 5e17])
    y = np.array([1.0, 2.0, np.nan, 4.0, np.nan, -np.inf, 0.5])
    for persp in ("global", "local"):
        ref = O.ici_kt(x, y, persp)
        st, out4, oc = run(epi, x, y, persp, "two.sided", False)
        assert st == ref.status == 0
        assert (oc[0], oc[1]) == (ref.xtie, ref.ytie)
        assert close(out4[0], ref.tau, 1e-12) and close(out4[1], ref.pvalue, 1e-9)


def test_epilogue_absorbed_minimum_joint_group(epi):
    # both columns absorb: the joint lowest group holds more rows than the joint-missing rows
    rng = np.random.default_rng(8)
    for trial in range(200):
        n = int(rng.integers(4, 60))
        x = np.round(rng.normal(size=n) * 2)
        y = np.round(rng.normal(size=n) * 2)
        x[rng.random(n) < 0.3] = -np.inf
        y[rng.random(n) < 0.3] = -np.inf if trial % 2 else -1e18
        x[rng.random(n) < 0.3] = np.nan
        y[rng.random(n) < 0.3] = np.nan
        for persp in ("global", "local"):
            ref = O.ici_kt(x, y, persp)
            st, out4, oc = run(epi, x, y, persp, "two.sided", False)
            assert st == ref.status
            if st == 0:
                assert (oc[0], oc[1], oc[2], oc[3]) == (ref.xtie, ref.ytie, ref.tot, ref.n_entry)
                assert close(out4[0], ref.tau, 1e-12) and close(out4[2], ref.tau_max, 1e-12)
                assert close(out4[1], ref.pvalue, 1e-9)


def test_epilogue_n2_gives_nan_pvalue(epi):
    x, y = np.array([1.0, 2.0]), np.array([2.0, 1.0])
    ref = O.ici_kt(x, y, "global")
    st, out4, _ = run(epi, x, y, "global", "two.sided", False)
    assert st == 0 and out4[0] == ref.tau == -1.0 and np.isnan(out4[1]) and np.isnan(ref.pvalue)


def test_completeness_reproduces_x87_long_double(epi):
    """src/kendallc.cpp:205-212 evaluates 1 - m/n in long double (x87 extended precision on the
    reference's x86-64 builds) and stores a double; the epilogue reproduces it bit for bit."""
    if np.finfo(np.longdouble).nmant != 63:
        pytest.skip("numpy longdouble is not the x87 80-bit format here")
    epi.one_minus_ratio_host.argtypes = [ctypes.c_ulonglong, ctypes.c_ulonglong]
    epi.one_minus_ratio_host.restype = ctypes.c_double
    rng = np.random.default_rng(5)
    one = np.longdouble(1)
    cases = [(m, n) for n in (1, 2, 3, 7, 10, 96, 100, 1000, 4099, 28672, 65535) for m in range(0, min(n, 300) + 1)]
    cases += [(n - d, n) for n in (5000, 20000, 28672, 60000, 65535) for d in range(0, 400)]
    cases += [(int(rng.integers(0, n + 1)), n) for n in rng.integers(1, 2 ** 31 - 1, size=20000).tolist()]
    bad = 0
    for m, n in cases:
        want = float(one - np.longdouble(m) / np.longdouble(n))
        bad += epi.one_minus_ratio_host(m, n) != want
    assert bad == 0
