"""The R .Call shim EXECUTED on the GPU box through the stand-in R C API (tests/r_stub/): every
registered routine is called the way r_package/R/icikt_b200.R calls it and its result compared
with the ctypes binding of the same library; degenerate pairs must come back as NA_real_ (payload
1954, what testthat's waldo tells apart from NaN, tests/testthat/test-kendall-tau.R:45-51), the
PROTECT stack must be balanced on success and on error, and a vector of device ordinals must take
the multi-GPU path."""
import numpy as np
import pytest

import icikendalltau_b200 as ik
from icikendalltau_b200 import _lib
from tests.r_stub import harness

pytestmark = pytest.mark.gpu


@pytest.fixture()
def sh():
    s = harness.load()
    s.reset()
    yield s
    assert s.protect_depth() == 0 and not s.protect_underflow()
    s.reset()


def _matrix(seed=3, n=700, C=9):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, 1)) + rng.normal(size=(n, C))
    x[:, 2] = np.round(x[:, 2] * 2)
    x = np.where(x < np.quantile(x, 0.2), np.nan, x)
    x[:, C - 4] = 4.25    # single unique value -> status 3
    x[:, C - 2] = np.nan  # all missing -> status 1
    return np.asfortranarray(x)


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("persp,diag", [("global", False), ("local", True)])
def test_all_pairs_entry_matches_binding(sh, persp, diag):
    x = _matrix()
    r = sh.dot_call("C_icikt_all_pairs", sh.real_matrix(x), sh.real([]), sh.string(persp), sh.string("two.sided"),
                    sh.logical(False), sh.logical(diag), sh.logical(False), sh.integer([0]))
    assert list(r) == ["raw", "pvalue", "taumax", "completeness", "status", "max_taumax"]
    ref = ik.run_pairs(x, (), perspective=persp, include_diag=diag)
    assert np.array_equal(r["status"], ref["status"]) and r["status"].dtype == np.int32
    bad = ref["status"] != 0
    assert bad.any() and (~bad).any()
    for k in ("raw", "pvalue", "taumax", "completeness"):
        assert _same(r[k][~bad], ref[k][~bad]), k
        assert sh.is_na(r[k][bad]).all(), f"{k}: degenerate pairs must be NA_real_, not NaN"
        assert not sh.is_na(r[k][~bad]).any()
    assert r["max_taumax"][0] == ref["max_taumax"]
    assert sh.protect_max() >= 1


def test_all_pairs_device_vector_takes_the_multi_gpu_path(sh):
    x = _matrix(seed=4, n=900, C=12)
    ndev = _lib.load().icikt_device_count()
    devs = [0, 1 % ndev, 0]
    one = sh.dot_call("C_icikt_all_pairs", sh.real_matrix(x), sh.real([0.0]), sh.string("global"), sh.string("less"),
                      sh.logical(True), sh.logical(False), sh.logical(True), sh.integer([0]))
    many = sh.dot_call("C_icikt_all_pairs", sh.real_matrix(x), sh.real([0.0]), sh.string("global"), sh.string("less"),
                       sh.logical(True), sh.logical(False), sh.logical(True), sh.integer(devs))
    for k in one:
        assert _same(one[k], many[k]), k
    ref = ik.run_pairs(x, (0.0, np.inf), perspective="global", alternative="less", continuity=True)
    ok = ref["status"] == 0
    assert _same(one["pvalue"][ok], ref["pvalue"][ok]) and _same(one["raw"][ok], ref["raw"][ok])


def test_pair_list_entry_is_one_based_and_covers_ici_kt(sh):
    x = _matrix(seed=5)
    pi = np.array([1, 3, 9, 6, 4], dtype=np.int32)  # R indices
    pj = np.array([2, 1, 9, 2, 8], dtype=np.int32)
    r = sh.dot_call("C_icikt_pair_list", sh.real_matrix(x), sh.real([]), sh.integer(pi), sh.integer(pj),
                    sh.string("local"), sh.string("greater"), sh.logical(False), sh.logical(False), sh.integer([0]))
    ref = ik.run_pairs(x, (), pi=pi - 1, pj=pj - 1, perspective="local", alternative="greater")
    assert np.array_equal(r["status"], ref["status"])
    ok = ref["status"] == 0
    for k in ("raw", "pvalue", "taumax", "completeness"):
        assert _same(r[k][ok], ref[k][ok]), k
        assert sh.is_na(r[k][~ok]).all()
    # ici_kt(x, y): cbind(x, y), i = 1L, j = 2L (r_package/R/icikt_b200.R ici_kt); n == 2 keeps its
    # genuine NaN p-value (status 0), which must NOT be rewritten to NA
    two = np.asfortranarray(np.array([[1.0, 2.0], [2.0, 1.0]]))
    r2 = sh.dot_call("C_icikt_pair_list", sh.real_matrix(two), sh.real([]), sh.integer([1]), sh.integer([2]),
                     sh.string("local"), sh.string("two.sided"), sh.logical(False), sh.logical(False), sh.integer([0]))
    assert r2["status"][0] == 0 and r2["raw"][0] == -1.0
    assert np.isnan(r2["pvalue"][0]) and not sh.is_na(r2["pvalue"])[0]
    with pytest.raises(RuntimeError, match=r"libicikt_b200 \(-2\)"):  # index 0 is not a column in R
        sh.dot_call("C_icikt_pair_list", sh.real_matrix(two), sh.real([]), sh.integer([0]), sh.integer([2]),
                    sh.string("local"), sh.string("two.sided"), sh.logical(False), sh.logical(False), sh.integer([0]))


@pytest.mark.parametrize("scale_max,diag_good,listed", [(True, True, False), (False, False, False), (True, True, True)])
def test_matrices_entry_matches_binding(sh, scale_max, diag_good, listed):
    x = _matrix(seed=6, n=500, C=8)
    n_good = (~np.isnan(x) & (x != 0)).sum(axis=0).astype(np.int32)
    pi = np.array([1, 1, 2, 4], dtype=np.int32) if listed else None
    pj = np.array([2, 4, 3, 7], dtype=np.int32) if listed else None
    r = sh.dot_call("C_icikt_matrices", sh.real_matrix(x), sh.real([0.0]),
                    sh.integer(pi) if listed else sh.nil(), sh.integer(pj) if listed else sh.nil(),
                    sh.string("global"), sh.string("two.sided"), sh.logical(False), sh.logical(True), sh.integer([0]),
                    sh.logical(scale_max), sh.logical(diag_good), sh.integer(n_good))
    assert list(r) == ["cor", "raw", "pvalue", "taumax", "completeness", "status_counts", "max_taumax"]
    ref = _lib.run_matrices(x, (0.0, np.inf), scale_max=scale_max, diag_good=diag_good, n_good=n_good,
                            pi=None if not listed else pi - 1, pj=None if not listed else pj - 1,
                            perspective="global")
    for k in ("cor", "raw", "pvalue", "taumax", "completeness"):
        assert r[k].shape == (8, 8)
        assert np.array_equal(r[k].view(np.uint64), ref[k].view(np.uint64)), k  # bit patterns: NA_real_ included
    assert np.array_equal(r["status_counts"].astype(np.int64), ref["status_counts"])
    assert sh.is_na(r["cor"][4, 0]) if not listed else True  # the constant column: NA, symmetric
    assert _same(r["raw"], r["raw"].T)
    if not listed:  # a vector of ordinals: icikt_matrices_multi, the very same bytes
        ndev = _lib.load().icikt_device_count()
        rm = sh.dot_call("C_icikt_matrices", sh.real_matrix(x), sh.real([0.0]), sh.nil(), sh.nil(),
                         sh.string("global"), sh.string("two.sided"), sh.logical(False), sh.logical(True),
                         sh.integer([0, 1 % ndev, 0]), sh.logical(scale_max), sh.logical(diag_good), sh.integer(n_good))
        for k in r:
            assert np.array_equal(r[k].view(np.uint64), rm[k].view(np.uint64)), k


def test_pairwise_completeness_entry_matches_binding(sh):
    x = _matrix(seed=8, n=333, C=7)
    gna = [np.nan, np.inf, 0.0]
    r = sh.dot_call("C_icikt_pairwise_completeness", sh.real_matrix(x), sh.real(gna), sh.nil(), sh.nil(),
                    sh.integer([0]), sh.logical(True))
    ref = _lib.pairwise_completeness(x, gna, want_matrix=True)
    assert np.array_equal(r["missingness"], ref["missing"]) and _same(r["completeness"], ref["completeness"])
    assert _same(r["matrix"], ref["matrix"])
    pi, pj = np.array([1, 7], dtype=np.int32), np.array([3, 7], dtype=np.int32)
    r2 = sh.dot_call("C_icikt_pairwise_completeness", sh.real_matrix(x), sh.real(gna), sh.integer(pi), sh.integer(pj),
                     sh.integer([0]), sh.logical(True))
    ref2 = _lib.pairwise_completeness(x, gna, pi=pi - 1, pj=pj - 1)
    assert np.array_equal(r2["missingness"], ref2["missing"]) and r2["matrix"] is None


def test_library_failure_is_an_r_error_with_balanced_protect(sh):
    long = np.zeros((65536, 2))  # n > icikt_max_n()
    long[:, 0] = np.arange(65536)
    with pytest.raises(RuntimeError, match=r"libicikt_b200 \(-3\): n exceeds icikt_max_n"):
        sh.dot_call("C_icikt_all_pairs", sh.real_matrix(long), sh.real([]), sh.string("global"), sh.string("two.sided"),
                    sh.logical(False), sh.logical(False), sh.logical(False), sh.integer([0]))
    assert sh.protect_depth() == 0 and not sh.protect_underflow() and sh.protect_max() == 1
    with pytest.raises(RuntimeError, match=r"libicikt_b200 \(-2\)"):  # a device ordinal that does not exist
        sh.dot_call("C_icikt_all_pairs", sh.real_matrix(long[:100]), sh.real([]), sh.string("global"),
                    sh.string("two.sided"), sh.logical(False), sh.logical(False), sh.logical(False), sh.integer([4096]))
