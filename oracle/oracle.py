"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes binding of the C++ restatement (oracle/icikt_oracle.cpp) plus a NumPy
restatement of the reference's R-level driver (paths relative to /root/reference):

  setup_missing_matrix   R/utils.R:1-23
  setup_comparisons      R/kendalltau.R:181-278   (pair order, include_only)
  ici_split              R/kendalltau.R:280-308   (pair loop)
  scale_and_reshape      R/kendalltau.R:357-421
  ici_kendalltau         R/kendalltau.R:96-179
  kt_split / kt_fast     R/kendalltau.R:310-354, 448-545
  pairwise_completeness  R/kendalltau.R:563-629

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Column "names" are 0-based integer indices here; R's
character names are a presentation detail of the R shim.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libicikt_oracle.so")


class _Result(ctypes.Structure):
    _fields_ = [
        ("tau", ctypes.c_double), ("pvalue", ctypes.c_double),
        ("tau_max", ctypes.c_double), ("completeness", ctypes.c_double),
        ("dis", ctypes.c_int64), ("ntie", ctypes.c_int64), ("xtie", ctypes.c_int64),
        ("ytie", ctypes.c_int64), ("tot", ctypes.c_int64), ("n_entry", ctypes.c_int64),
        ("x0", ctypes.c_int64), ("y0", ctypes.c_int64), ("x1", ctypes.c_int64),
        ("y1", ctypes.c_int64), ("n_matching_na", ctypes.c_int64),
        ("z", ctypes.c_double), ("var", ctypes.c_double),
        ("status", ctypes.c_int32),
    ]


_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (g++ -O2)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "icikt_oracle.cpp"))):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int32)
        lp = ctypes.POINTER(ctypes.c_int64)
        L.icikt_oracle_ici_kt.argtypes = [dp, ctypes.c_int64, dp, ctypes.c_int64, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.POINTER(_Result)]
        L.icikt_oracle_ici_kt.restype = ctypes.c_int
        L.icikt_oracle_ici_kt_pairs.argtypes = [dp, ctypes.c_int64, dp, ctypes.c_int64,
                                                ctypes.c_int, ctypes.c_int, dp]
        L.icikt_oracle_ici_kt_pairs.restype = ctypes.c_int
        L.icikt_oracle_pnorm.argtypes = [ctypes.c_double, ctypes.c_int]
        L.icikt_oracle_pnorm.restype = ctypes.c_double
        L.icikt_oracle_pair_loop.argtypes = [dp, ctypes.c_int64, ctypes.c_int64, ip, ip,
                                             ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             dp, dp, dp, dp, ip, lp]
        L.icikt_oracle_pair_loop.restype = ctypes.c_int
        L.icikt_oracle_pair_loop_z.argtypes = L.icikt_oracle_pair_loop.argtypes + [dp]
        L.icikt_oracle_pair_loop_z.restype = ctypes.c_int
        _lib = L
    return _lib


PERSPECTIVE = {"global": 0, "local": 1}
ALTERNATIVE = {"two.sided": 0, "less": 1, "greater": 2}
WARNINGS = {
    2: "Warning: The vectors only have a single value, NA returned!",
    3: "Warning: Either 'X' or 'Y' have only a single unique value, NA returned!",
    4: "Warning: Ties equal the total, NA returned!",
}


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


@dataclass
class PairResult:
    tau: float
    pvalue: float
    tau_max: float
    completeness: float
    status: int
    dis: int
    ntie: int
    xtie: int
    ytie: int
    tot: int
    n_entry: int
    x0: int
    y0: int
    x1: int
    y1: int
    n_matching_na: int
    z: float
    var: float

    def as_vector(self):
        return np.array([self.tau, self.pvalue, self.tau_max, self.completeness])


def ici_kt(x, y, perspective="local", alternative="two.sided", continuity=False,
           emulate_int32=False) -> PairResult:
    """ici_kt, src/kendallc.cpp:166-366 (defaults as R/RcppExports.R:62)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    r = _Result()
    rc = lib().icikt_oracle_ici_kt(_dptr(x), x.size, _dptr(y), y.size,
                                   PERSPECTIVE.get(perspective, 0),
                                   ALTERNATIVE.get(alternative, 3), int(bool(continuity)),
                                   int(bool(emulate_int32)), ctypes.byref(r))
    if rc == -1:
        raise ValueError("'X' and 'Y' are not the same length!")  # src/kendallc.cpp:168-170
    return PairResult(*[getattr(r, f) for f in
                        ("tau", "pvalue", "tau_max", "completeness", "status", "dis", "ntie",
                         "xtie", "ytie", "tot", "n_entry", "x0", "y0", "x1", "y1",
                         "n_matching_na", "z", "var")])


def ici_kt_pairs(x, y, perspective="local", alternative="two.sided"):
    """ici_kt_pairs, src/kendallc.cpp:370-549.  Returns (tau, pvalue)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.zeros(2)
    rc = lib().icikt_oracle_ici_kt_pairs(_dptr(x), x.size, _dptr(y), y.size,
                                         PERSPECTIVE.get(perspective, 0),
                                         ALTERNATIVE.get(alternative, 3), _dptr(out))
    if rc == -1:
        raise ValueError("X and Y are not the same length!")
    return out


def pnorm(x, lower_tail=True):
    return lib().icikt_oracle_pnorm(float(x), int(bool(lower_tail)))


# ----------------------------------------------------------------------------------
# R-level driver
# ----------------------------------------------------------------------------------

def setup_missing_matrix(data, global_na=(np.nan, np.inf, 0)):
    """R/utils.R:1-23."""
    data = np.asarray(data, dtype=np.float64)
    excl = np.zeros(data.shape, dtype=bool)
    g = [float(v) for v in global_na]
    if len(g) > 0:
        if any(np.isnan(v) for v in g):
            excl |= np.isnan(data)
            g = [v for v in g if not np.isnan(v)]
        if any(np.isinf(v) for v in g):
            excl |= np.isinf(data)
            g = [v for v in g if not np.isinf(v)]
    for v in g:
        excl |= (data == v)
    return excl


def setup_comparisons(n_sample, include_only=None, diag_good=True):
    """R/kendalltau.R:181-247: combn order, (i,i) appended when !diag_good,
    include_only filtering.  Returns (pi, pj) 0-based int32 arrays."""
    iu = np.triu_indices(n_sample, k=1)  # row-major upper triangle == utils::combn order
    pi, pj = iu[0].astype(np.int32), iu[1].astype(np.int32)
    if not diag_good:
        d = np.arange(n_sample, dtype=np.int32)
        pi, pj = np.concatenate([pi, d]), np.concatenate([pj, d])
    if include_only is not None:
        if isinstance(include_only, (list, tuple)) and len(include_only) > 0 and \
                isinstance(include_only[0], (list, tuple, np.ndarray)):
            if len(include_only) != 2:
                raise ValueError("must be a vector, a data.frame with two columns, or list of two vectors.")
            l1 = np.atleast_1d(np.asarray(include_only[0]))
            l2 = np.atleast_1d(np.asarray(include_only[1]))
            l1, l2 = np.broadcast_arrays(l1, l2)  # paste0 recycles
            want = set((int(a), int(b)) for a, b in zip(l1, l2)) | \
                set((int(b), int(a)) for a, b in zip(l1, l2))
            keep = np.array([(int(a), int(b)) in want for a, b in zip(pi, pj)], dtype=bool)
        else:
            inc = np.atleast_1d(np.asarray(include_only))
            keep = np.isin(pi, inc) | np.isin(pj, inc)
        pi, pj = pi[keep], pj[keep]
    if pi.size == 0:
        raise ValueError("No comparisons to do.")
    return pi, pj


def pair_loop(exclude_data, pi, pj, perspective="global", alternative="two.sided",
              continuity=False, ncore=1, emulate_int32=False, want_counts=False, want_z=False):
    """ici_split over split_comparisons, R/kendalltau.R:158,280-308.
    want_z: also return the normal deviate z of every pair (src/kendallc.cpp:321)."""
    data = np.asfortranarray(exclude_data, dtype=np.float64)
    n, C = data.shape
    pi = np.ascontiguousarray(pi, dtype=np.int32)
    pj = np.ascontiguousarray(pj, dtype=np.int32)
    P = pi.size
    raw, pv, tm, comp = (np.empty(P) for _ in range(4))
    status = np.empty(P, dtype=np.int32)
    counts = np.empty((P, 7), dtype=np.int64) if want_counts else None
    ip = ctypes.POINTER(ctypes.c_int32)
    z = np.empty(P) if want_z else None
    lib().icikt_oracle_pair_loop_z(
        _dptr(data), n, C, pi.ctypes.data_as(ip), pj.ctypes.data_as(ip), P,
        PERSPECTIVE.get(perspective, 0), ALTERNATIVE.get(alternative, 3), int(bool(continuity)),
        int(bool(emulate_int32)), int(ncore), _dptr(raw), _dptr(pv), _dptr(tm), _dptr(comp),
        status.ctypes.data_as(ip),
        counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)) if want_counts else None,
        _dptr(z) if want_z else None)
    out = dict(raw=raw, pvalue=pv, taumax=tm, completeness=comp, status=status)
    if want_z:
        out["z"] = z
    if want_counts:
        out["counts"] = counts
    return out


def ici_kendalltau(data_matrix, global_na=(np.nan, np.inf, 0), perspective="global",
                   scale_max=True, diag_good=True, include_only=None, alternative="two.sided",
                   continuity=False, return_matrix=True, ncore=1):
    """R/kendalltau.R:96-179 + scale_and_reshape :357-421."""
    data = np.asarray(data_matrix, dtype=np.float64)
    excl = setup_missing_matrix(data, global_na)
    ex = data.copy()
    ex[excl] = np.nan
    n, C = ex.shape
    pi, pj = setup_comparisons(C, include_only, diag_good)
    res = pair_loop(ex, pi, pj, perspective, alternative, continuity, ncore)
    raw, pv, tm, comp = res["raw"], res["pvalue"], res["taumax"], res["completeness"]
    if scale_max:
        max_cor = np.nanmax(tm)  # max(taumax, na.rm = TRUE), :368-370
        cor = raw / max_cor
    else:
        cor = raw.copy()
    n_good = (~excl).sum(axis=0)
    frac_complete = n_good / n
    if diag_good:  # :374-386, appended after scaling
        d = np.arange(C, dtype=np.int32)
        pi, pj = np.concatenate([pi, d]), np.concatenate([pj, d])
        dg = n_good / n_good.max()
        raw = np.concatenate([raw, dg])
        cor = np.concatenate([cor, dg])
        pv = np.concatenate([pv, np.zeros(C)])
        tm = np.concatenate([tm, np.ones(C)])
        comp = np.concatenate([comp, frac_complete])
    if not return_matrix:
        return dict(s1=pi, s2=pj, raw=raw, pvalue=pv, taumax=tm, completeness=comp, cor=cor)
    out = {}
    for name, v in (("cor", cor), ("raw", raw), ("pvalue", pv), ("taumax", tm), ("completeness", comp)):
        m = np.zeros((C, C))
        m[pi, pj] = v
        m[pj, pi] = v
        out[name] = m
    out["keep"] = (~excl).T
    return out


def kt_fast(x, use="everything", return_matrix=True):
    """R/kendalltau.R:448-545 with kt_split :310-354, for a matrix x (columns = variables).
    alternative/continuity are accepted by the reference but never forwarded (:341)."""
    x = np.asarray(x, dtype=np.float64)
    if use == "na.or.complete":
        raise ValueError("'na.or.complete' is not a supported value for use.")
    na = np.isnan(x)
    any_na = na.any()
    no_na_rows = na.sum(axis=1) == 0
    C = x.shape[1]
    pi, pj = setup_comparisons(C, None, diag_good=False)
    P = pi.size
    do = True
    if use in ("everything", "all.obs") and any_na:
        do = False
    if use in ("complete.obs",):  # :491 tests the misspelt "pariwise.complete.obs"
        if no_na_rows.sum() == 0:
            do = False
        else:
            x = x[no_na_rows, :]
    tau = np.full(P, np.nan)
    pv = np.full(P, np.nan)
    if do:
        for k in range(P):
            tx, ty = x[:, pi[k]], x[:, pj[k]]
            ret_na = False
            if use == "pairwise.complete.obs":
                good = ~np.isnan(tx) & ~np.isnan(ty)
                if good.sum() == 0:
                    ret_na = True
                else:
                    tx, ty = tx[good], ty[good]
            elif use in ("everything", "all.obs"):
                ret_na = bool(np.isnan(tx).any() or np.isnan(ty).any())
            if not ret_na:
                r = ici_kt(tx, ty)  # defaults: local, two.sided, no continuity
                tau[k], pv[k] = r.tau, r.pvalue
    if not return_matrix:
        return dict(s1=pi, s2=pj, tau=tau, pvalue=pv)
    tm = np.zeros((C, C))
    pm = np.zeros((C, C))
    tm[pi, pj] = tau
    tm[pj, pi] = tau
    pm[pi, pj] = pv
    pm[pj, pi] = pv
    return dict(tau=tm, pvalue=pm)


def pairwise_completeness(data_matrix, global_na=(np.nan, np.inf, 0), include_only=None,
                          return_matrix=True):
    """R/kendalltau.R:563-629."""
    data = np.asarray(data_matrix, dtype=np.float64)
    excl = setup_missing_matrix(data, global_na)
    n, C = excl.shape
    pi, pj = setup_comparisons(C, include_only, diag_good=False)
    missing = (excl[:, pi] | excl[:, pj]).sum(axis=0)
    comp = 1 - missing / n
    if not return_matrix:
        return dict(s1=pi, s2=pj, missingness=missing, completeness=comp)
    m = np.zeros((C, C))
    m[pi, pj] = comp
    m[pj, pi] = comp
    return m
