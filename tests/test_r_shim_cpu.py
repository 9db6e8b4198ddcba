"""The R .Call shim (r_package/src/icikt_shim.c) compiled and EXECUTED against a stand-in of R's C
API (tests/r_stub/: no R in this image).  CPU part: it builds warning-free with -Werror, registers
its routines like src/RcppExports.cpp:113-128 of the reference does, the arities in CallEntries[]
equal what r_package/R/icikt_b200.R passes to .Call, argument errors surface as R errors with a
balanced PROTECT stack, and -- without a GPU -- the library's "no CUDA device" failure comes back
as an R error (there is no CPU fallback to fall into)."""
import os
import re

import numpy as np
import pytest

from icikendalltau_b200 import _lib
from tests.r_stub import harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXPECTED = {"C_icikt_all_pairs": 8, "C_icikt_pair_list": 9, "C_icikt_matrices": 12,
            "C_icikt_pairwise_completeness": 6, "C_icikt_device_count": 0, "C_icikt_release": 0}


@pytest.fixture()
def sh():
    s = harness.load()
    s.reset()
    yield s
    s.reset()


def test_registration_table(sh):
    assert sh.routines() == EXPECTED
    assert sh.use_dynamic_symbols() is False  # R_useDynamicSymbols(dll, FALSE), src/RcppExports.cpp:127


def _call_sites(text):
    """(.Call symbol, number of arguments) of every .Call( in the R source (balanced-parenthesis scan)."""
    out = []
    for m in re.finditer(r"\.Call\((C_[A-Za-z_]+)", text):
        depth, k, args, cur = 1, m.end(), 0, ""
        while depth:
            c = text[k]
            if c in "([":
                depth += 1
            elif c in ")]":
                depth -= 1
            if depth == 1 and c == "," or depth == 0:
                if cur.strip(", \n"):
                    args += 1
                cur = ""
            else:
                cur += c
            k += 1
        out.append((m.group(1), args - 0))
    return out


def test_r_call_sites_match_registered_arities(sh):
    text = open(os.path.join(ROOT, "r_package", "R", "icikt_b200.R")).read()
    text = "\n".join(l.split("#")[0] for l in text.splitlines())  # drop comments
    sites = _call_sites(text)
    assert len(sites) >= 8
    seen = set()
    for name, nargs in sites:
        # the symbol itself is the first thing inside .Call(...): nargs counts what follows it
        assert EXPECTED[name] == nargs, f".Call({name}, ...) passes {nargs} arguments, CallEntries[] says {EXPECTED[name]}"
        seen.add(name)
    assert seen == set(EXPECTED)
    # and .Call refuses a wrong count before reaching C, like R
    with pytest.raises(RuntimeError, match="Incorrect number of arguments"):
        sh.dot_call("C_icikt_release", sh.nil())
    with pytest.raises(RuntimeError, match="not in load table"):
        sh.dot_call("_ICIKendallTau_ici_kt")


def test_all_pair_plan_never_builds_combn():
    """the R host must not build utils::combn(5000, 2) for the all-pairs matrix path."""
    text = open(os.path.join(ROOT, "r_package", "R", "icikt_b200.R")).read()
    code = "\n".join(l.split("#")[0] for l in text.splitlines())
    assert "utils::combn(" not in code and "combn(" not in code
    assert "need_indices = !on_device || check_timing" in code


def _args_all_pairs(sh, x, persp="global", device=0):
    dev = sh.integer([device] if np.isscalar(device) else device)
    return (sh.real_matrix(x), sh.real([]), sh.string(persp), sh.string("two.sided"), sh.logical(False),
            sh.logical(False), sh.logical(False), dev)


def test_argument_errors_are_r_errors_with_balanced_protect(sh):
    x = np.random.default_rng(0).normal(size=(30, 4))
    bad = list(_args_all_pairs(sh, x))
    bad[0] = sh.int_matrix(np.ones((3, 3)))  # integer matrix: the R host converts, the shim insists
    with pytest.raises(RuntimeError, match="`data` must be a double matrix"):
        sh.dot_call("C_icikt_all_pairs", *bad)
    bad = list(_args_all_pairs(sh, x))
    bad[0] = sh.real(x[:, 0])  # a plain vector is not a matrix
    with pytest.raises(RuntimeError, match="`data` must be a double matrix"):
        sh.dot_call("C_icikt_all_pairs", *bad)
    bad = list(_args_all_pairs(sh, x))
    bad[1] = sh.integer([0])
    with pytest.raises(RuntimeError, match="`global_na` must be a double vector"):
        sh.dot_call("C_icikt_all_pairs", *bad)
    with pytest.raises(RuntimeError, match="integer vectors of one length"):
        sh.dot_call("C_icikt_pair_list", sh.real_matrix(x), sh.real([]), sh.integer([1, 2]), sh.integer([2]),
                    sh.string("local"), sh.string("two.sided"), sh.logical(False), sh.logical(False), sh.integer([0]))
    with pytest.raises(RuntimeError, match="one entry per column"):
        sh.dot_call("C_icikt_matrices", sh.real_matrix(x), sh.real([]), sh.nil(), sh.nil(), sh.string("global"),
                    sh.string("two.sided"), sh.logical(False), sh.logical(False), sh.integer([0]), sh.logical(True),
                    sh.logical(True), sh.integer([30, 30]))
    assert sh.protect_depth() == 0 and not sh.protect_underflow()


def test_device_count_and_release_entries(sh):
    n = sh.dot_call("C_icikt_device_count")
    assert n.shape == (1,) and n[0] == _lib.load().icikt_device_count()
    assert sh.dot_call("C_icikt_release") is None
    assert sh.protect_depth() == 0


@pytest.mark.skipif(_lib.load().icikt_device_count() > 0, reason="checks the no-GPU behaviour")
def test_no_device_is_an_r_error_not_a_fallback(sh):
    x = np.random.default_rng(1).normal(size=(40, 5))
    for name, args in (
            ("C_icikt_all_pairs", _args_all_pairs(sh, x)),
            ("C_icikt_pair_list", (sh.real_matrix(x), sh.real([]), sh.integer([1]), sh.integer([2]), sh.string("local"),
                                   sh.string("two.sided"), sh.logical(False), sh.logical(False), sh.integer([0]))),
            ("C_icikt_matrices", (sh.real_matrix(x), sh.real([0.0]), sh.nil(), sh.nil(), sh.string("global"),
                                  sh.string("two.sided"), sh.logical(False), sh.logical(True), sh.integer([0]),
                                  sh.logical(True), sh.logical(True), sh.integer([40] * 5))),
            ("C_icikt_pairwise_completeness", (sh.real_matrix(x), sh.real([np.nan, 0.0]), sh.nil(), sh.nil(),
                                               sh.integer([0]), sh.logical(True)))):
        with pytest.raises(RuntimeError, match=r"libicikt_b200 \(-1\): no CUDA device"):
            sh.dot_call(name, *args)
        # the shim drops its PROTECTs before error(): nothing is left on the stack, nothing underflowed
        assert sh.protect_depth() == 0 and not sh.protect_underflow() and sh.protect_max() >= 1
