"""CPU-only checks of the drop-in boundary: the shared library builds/loads, exports every
symbol include/icikt_b200.h declares, refuses to compute without a GPU (no CPU fallback), and
the host-side planning logic matches the oracle's restatement of the R code."""
import ctypes
import os
import re

import numpy as np
import pytest

import icikendalltau_b200 as ik
from icikendalltau_b200 import _lib, api
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        return _lib.load().icikt_device_count() > 0
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "icikt_b200.h")).read()
    declared = set(re.findall(r"\b(icikt_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"icikt_max_n"} - {"icikt_max_n"}  # keep all
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(_lib.EXPORTS) <= declared
    assert _lib.load().icikt_abi_version() == 1
    assert _lib.load().icikt_max_n() >= 20000


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.Opts) == 8 * 4 + 2 * 8
    assert ctypes.sizeof(_lib.Timings) == 6 * 4 + 2 * 4
    assert ctypes.sizeof(_lib.Table) == 16 and _lib.MAX_TABLES == 16  # icikt_table, ICIKT_MAX_TABLES


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    x = np.random.default_rng(0).normal(size=(20, 3))
    with pytest.raises(ik.IciktError) as e:
        ik.run_pairs(x)
    assert e.value.code == _lib.ERR_NO_DEVICE
    with pytest.raises(ik.IciktError):
        ik.ici_kt(x[:, 0], x[:, 1])
    with pytest.raises(ik.IciktError):
        ik.Plan(20, 3)
    # a job large enough for the pipelined one-shot call (>= 32 MB, >= 128 columns) fails the same way, before
    # any stream or staging buffer exists
    big = np.zeros((16500, 256), order="F")
    with pytest.raises(ik.IciktError) as e:
        ik.run_pairs(big)
    assert e.value.code == _lib.ERR_NO_DEVICE
    with pytest.raises(ik.IciktError) as e:
        _lib.run_matrices(big, want=("raw",))
    assert e.value.code == _lib.ERR_NO_DEVICE


def test_argument_errors_before_any_device_work():
    x = np.arange(10.0)
    with pytest.raises(ValueError, match="not the same length"):  # src/kendallc.cpp:168-170
        ik.ici_kt(x, x[:9])
    m = np.random.default_rng(0).normal(size=(20, 10))
    with pytest.raises(ValueError, match="Colnames of `data_matrix` must be be specified."):
        ik.ici_kendalltau(m)
    with pytest.raises(ValueError, match="Colnames of `x` must be be specified."):
        ik.kt_fast(m)
    with pytest.raises(TypeError, match="must be a numeric type"):
        ik.ici_kendalltau(m.astype(str), colnames=[f"S{i}" for i in range(10)])
    with pytest.raises(ValueError, match="is not a supported"):
        ik.kt_fast(m, use="na.or.complete", colnames=list("abcdefghij"))
    with pytest.raises(ValueError, match="should both be provided as vectors"):
        ik.kt_fast(m[:, 0])
    with pytest.raises(ValueError, match="must be vectors"):
        ik.kt_fast(m, m)


def test_setup_comparisons_matches_oracle():
    names = [f"s{i + 1}" for i in range(100)]
    for inc_names, inc_idx, diag in (
            (None, None, True), (None, None, False),
            ("s1", 0, True), (["s1", "s3"], [0, 2], True),
            ((["s1"], ["s2", "s3"]), ([0], [1, 2]), True),
            ((["s1"], ["s2", "s3"]), ([0], [1, 2]), False)):
        pi, pj, allp = api.setup_comparisons(names, inc_names, diag)
        opi, opj = O.setup_comparisons(100, inc_idx, diag)
        assert np.array_equal(pi, opi) and np.array_equal(pj, opj)
        assert allp == (inc_names is None)
    # the documented list(g1 = "s1", g2 = c("s2", "s3")) form (R/kendalltau.R:86-91): a scalar member is
    # a vector of length one, as a list, a tuple or a dict
    want = api.setup_comparisons(names, (["s1"], ["s2", "s3"]), True)
    for form in (["s1", ["s2", "s3"]], ("s1", ("s2", "s3")), {"g1": "s1", "g2": ["s2", "s3"]},
                 [np.array(["s2", "s3"]), "s1"]):
        got = api.setup_comparisons(names, form, True)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and got[2] is False
    # test-kendall-tau.R:102-136: zero counts of the C x C result follow from the pair counts
    for inc, zeros in (("s1", 9702), (["s1", "s3"], 9506), ((["s1"], ["s2", "s3"]), 9896)):
        pi, pj, _ = api.setup_comparisons(names, inc, True)
        n_off = pi.size
        assert 100 * 100 - 2 * n_off - 100 == zeros
    pi, pj, _ = api.setup_comparisons(names, (["s1"], ["s2", "s3"]), False)
    assert pi.size == 2 and 100 * 100 - 2 * pi.size == 9996
    with pytest.raises(ValueError, match="list of two vectors"):
        api.setup_comparisons(names, (["s1"], ["s2", "s3"], ["s4"]), False)
    with pytest.raises(ValueError, match="No comparisons to do."):
        api.setup_comparisons(names, (["s102"], ["s105"]), False)
    # check_timing structure (test-kendall-tau.R:256-265): 40 columns -> 780 pairs
    pi, _, _ = api.setup_comparisons([f"s{i}" for i in range(40)], None, True)
    assert pi.size == 780


def test_setup_missing_matrix_matches_oracle():
    rng = np.random.default_rng(1)
    x = rng.normal(size=(30, 5))
    x[rng.random(x.shape) < 0.1] = 0.0
    x[rng.random(x.shape) < 0.1] = np.nan
    x[3, 2] = np.inf
    x[4, 2] = -np.inf
    x[5, 1] = -2.0
    for g in ((np.nan, np.inf, 0), (np.nan,), (0,), (np.nan, np.inf, -2), ()):
        assert np.array_equal(api.setup_missing_matrix(x, g), O.setup_missing_matrix(x, g))


def test_run_pairs_result_views_match_the_addresses_handed_to_the_library(monkeypatch):
    """run_pairs hands the library plain addresses into ONE result allocation; the arrays it
    returns must be exactly those ranges (checked with a stand-in that writes through them)."""
    import ctypes
    real = _lib.load()
    seen = {}

    def fake_all_pairs(data, n, C, ld, gna, ngna, opts, raw, pv, tm, comp, status, counts, mx, timings):
        P = C * (C - 1) // 2
        seen.update(n=n, C=C, ld=ld, ngna=ngna, gna=gna)
        for k, addr in enumerate((raw, pv, tm, comp)):
            (ctypes.c_double * P).from_address(addr)[:] = [100.0 * k + i for i in range(P)]
        (ctypes.c_int32 * P).from_address(status)[:] = list(range(P))
        assert counts is None
        return 0

    class Stub:
        icikt_all_pairs = staticmethod(fake_all_pairs)

        def __getattr__(self, name):
            return getattr(real, name)

    monkeypatch.setattr(_lib, "_lib", Stub())
    x = np.asfortranarray(np.arange(35, dtype=np.float64).reshape(5, 7))
    r = _lib.run_pairs(x, (), perspective="local")
    P = 21
    for k, name in enumerate(("raw", "pvalue", "taumax", "completeness")):
        assert r[name].dtype == np.float64 and r[name].shape == (P,)
        assert np.array_equal(r[name], 100.0 * k + np.arange(P)), name
    assert r["status"].dtype == np.int32 and np.array_equal(r["status"], np.arange(P))
    assert (seen["n"], seen["C"], seen["ld"], seen["ngna"], seen["gna"]) == (5, 7, 5, 0, None)
    r["status"][3] = 0  # the views are writable (kt_fast clears status 9 in place)


def test_ici_kendalltau_matrix_path_host_logic(monkeypatch):
    """ici_kendalltau(return_matrix=True) hands scale_and_reshape to icikt_matrices: what the host
    layer passes (n_good of setup_missing_matrix, diag_good, scale_max, the include_only pair list)
    and what it does with the answer (warnings from the status counts, keep, names)."""
    import ctypes
    import warnings
    real = _lib.load()
    calls = []

    def fake_matrices(data, n, C, ld, gna, ngna, pi, pj, P, opts, scale_max, diag_good, n_good, cor, raw, pv, tm,
                      comp, hist, mx, timings):
        ng = list((ctypes.c_int32 * C).from_address(n_good)) if n_good else None
        pairs = None
        if pi:
            pairs = list(zip((ctypes.c_int32 * P).from_address(pi), (ctypes.c_int32 * P).from_address(pj)))
        calls.append(dict(n=n, C=C, scale_max=scale_max, diag_good=diag_good, n_good=ng, pairs=pairs,
                          literals=list((ctypes.c_double * ngna).from_address(gna)) if ngna else []))
        for k, addr in enumerate((cor, raw, pv, tm, comp)):
            (ctypes.c_double * (C * C)).from_address(addr)[:] = [float(k)] * (C * C)
        counts = (ctypes.c_int64 * _lib.NSTATUS).from_address(hist)
        counts[0], counts[3] = C * (C - 1) // 2 - 2, 2  # two pairs with a single unique value
        return 0

    class Stub:
        icikt_matrices = staticmethod(fake_matrices)

        def __getattr__(self, name):
            return getattr(real, name)

    monkeypatch.setattr(_lib, "_lib", Stub())
    x = np.array([[1.0, 0.0, 3.0], [2.0, 5.0, np.nan], [0.0, 6.0, 7.0], [4.0, np.inf, 8.0]])
    names = ["a", "b", "c"]
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        out = ik.ici_kendalltau(x, colnames=names, scale_max=False, diag_good=True)
    assert any("single unique value" in str(v.message) and "2 pairs" in str(v.message) for v in w)
    c = calls[-1]
    assert (c["n"], c["C"], c["scale_max"], c["diag_good"], c["pairs"]) == (4, 3, 0, 1, None)
    assert c["n_good"] == [3, 2, 3]  # zeros, NaN and Inf are missing by default (R/utils.R:1-23)
    assert sorted(v for v in c["literals"] if v == v and abs(v) != np.inf) == [0.0]
    assert out["raw"].shape == (3, 3) and np.all(out["cor"] == 0.0) and np.all(out["completeness"] == 4.0)
    assert out["keep"].shape == (3, 4) and out["keep"].sum() == 8 and out["names"] == names
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ik.ici_kendalltau(x, colnames=names, include_only="b", diag_good=False)
    c = calls[-1]
    assert c["diag_good"] == 0 and sorted(c["pairs"]) == [(0, 1), (1, 1), (1, 2)]


def test_kt_fast_pairwise_host_logic(monkeypatch):
    """kt_fast(use = "pairwise.complete.obs") on data with missing values: one all-pairs call in
    the complete-observations mode; pairs the device flags with status 9 are filtered on the host
    (R/kendalltau.R:323-331) and re-run as two-column problems."""
    import ctypes
    real = _lib.load()
    log = []

    def fake_all_pairs(data, n, C, ld, gna, ngna, opts, raw, pv, tm, comp, status, counts, mx, timings):
        o = ctypes.cast(opts, ctypes.POINTER(_lib.Opts)).contents
        P = C * (C - 1) // 2 + (C if o.include_diag else 0)
        log.append(("all", n, C, o.perspective, o.include_diag, P))
        (ctypes.c_double * P).from_address(raw)[:] = [0.5] * P
        (ctypes.c_double * P).from_address(pv)[:] = [0.25] * P
        st = (ctypes.c_int32 * P).from_address(status)
        st[:] = [0] * P
        st[1] = 9  # pair (0, 2): not supported on the device
        return 0

    def fake_pair_list(data, n, C, ld, gna, ngna, pi, pj, P, opts, raw, pv, tm, comp, status, counts, mx, timings):
        o = ctypes.cast(opts, ctypes.POINTER(_lib.Opts)).contents
        rows = np.ctypeslib.as_array((ctypes.c_double * (n * C)).from_address(data)).reshape(C, n).T.copy()
        log.append(("list", n, C, o.perspective, rows))
        (ctypes.c_double * P).from_address(raw)[:] = [-1.0] * P
        (ctypes.c_double * P).from_address(pv)[:] = [0.75] * P
        (ctypes.c_int32 * P).from_address(status)[:] = [0] * P
        return 0

    class Stub:
        icikt_all_pairs = staticmethod(fake_all_pairs)
        icikt_pair_list = staticmethod(fake_pair_list)

        def __getattr__(self, name):
            return getattr(real, name)

    monkeypatch.setattr(_lib, "_lib", Stub())
    x = np.array([[1.0, 4.0, 9.0], [2.0, np.nan, 8.0], [3.0, 6.0, np.nan], [5.0, 7.0, 6.0], [np.nan, 1.0, 5.0]])
    out = ik.kt_fast(x, use="pairwise.complete.obs", colnames=["a", "b", "c"], return_matrix=False)
    assert log[0] == ("all", 5, 3, _lib.PERSPECTIVE["complete"], 1, 6)
    kind, n, C, persp, rows = log[1]
    assert (kind, n, C, persp) == ("list", 3, 2, _lib.PERSPECTIVE["local"])  # rows present in both a and c
    assert np.array_equal(rows, np.array([[1.0, 9.0], [2.0, 8.0], [5.0, 6.0]]))
    tau = out["tau"]
    assert list(tau["s1"]) == ["a", "a", "b", "a", "b", "c"] and list(tau["s2"]) == ["b", "c", "c", "a", "b", "c"]
    assert list(tau["tau"]) == [0.5, -1.0, 0.5, 0.5, 0.5, 0.5] and tau["pvalue"][1] == 0.75
    # no missing value: the plain local perspective is enough
    log.clear()
    ik.kt_fast(np.nan_to_num(x, nan=0.5), use="pairwise.complete.obs", colnames=["a", "b", "c"])
    assert log[0][3] == _lib.PERSPECTIVE["local"] and len(log) == 1
