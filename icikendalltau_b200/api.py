"""Host-side mirror of the reference's R interface for the ICI-Kendall-tau path.

Same function names, argument meaning, result names and error/warning texts as the
reference (paths relative to the reference repository):

  ici_kt                 R/RcppExports.R:62-64 -> src/kendallc.cpp:166-366
  ici_kendalltau         R/kendalltau.R:96-179  (+ setup_comparisons :181-278,
                                                  scale_and_reshape :357-421)
  kt_fast                R/kendalltau.R:448-545 (+ kt_split :310-354)
  pairwise_completeness  R/kendalltau.R:563-629

All pair arithmetic runs in libicikt_b200.so on the GPU (no CPU fallback); this module
only validates arguments, plans the pair list and reshapes the results, which is what
the R host code keeps doing in the reference-side integration.
R is not available in the build image, so this Python layer is the executable host.
"""
from __future__ import annotations

import logging
import time
import warnings

import numpy as np

from . import _lib

log = logging.getLogger("icikendalltau_b200")

WARN_SINGLE_VALUE = "Warning: The vectors only have a single value, NA returned!"              # src/kendallc.cpp:225
WARN_SINGLE_UNIQUE = "Warning: Either 'X' or 'Y' have only a single unique value, NA returned!"  # :238
WARN_ALL_TIED = "Warning: Ties equal the total, NA returned!"                                   # :292
_STATUS_WARNING = {2: WARN_SINGLE_VALUE, 3: WARN_SINGLE_UNIQUE, 4: WARN_ALL_TIED}


class IciKtResult(dict):
    """Named vector c(tau, pvalue, tau_max, completeness) (src/kendallc.cpp:171-172)."""

    NAMES = ("tau", "pvalue", "tau_max", "completeness")

    def __getitem__(self, k):
        if isinstance(k, int):
            k = self.NAMES[k]
        return dict.__getitem__(self, k)

    def as_array(self):
        return np.array([dict.__getitem__(self, k) for k in self.NAMES])


def _warn_status(status):
    for code, text in _STATUS_WARNING.items():
        cnt = int((status == code).sum())
        if cnt:
            warnings.warn(text if cnt == 1 else f"{text} ({cnt} pairs)", RuntimeWarning, stacklevel=3)


def _warn_status_counts(counts):
    for code, text in _STATUS_WARNING.items():
        cnt = int(counts[code])
        if cnt:
            warnings.warn(text if cnt == 1 else f"{text} ({cnt} pairs)", RuntimeWarning, stacklevel=3)


def _colnames_of(x, colnames, arg):
    """check_if_colnames_null + transform_to_matrix + check_if_numeric (R/utils.R:25-66)."""
    names = colnames
    try:
        import pandas as pd
        if isinstance(x, pd.DataFrame):
            log.info("`%s` is a data.frame, converting to matrix ...", arg)
            names = list(x.columns) if names is None else names
            x = x.to_numpy()
    except ImportError:  # pandas is optional
        pass
    arr = np.asarray(x)
    if arr.ndim != 2:
        raise ValueError(f"`{arg}` must be matrix-like (features x samples)")
    if names is None:
        raise ValueError(f"Colnames of `{arg}` must be be specified.")
    names = list(names)
    if len(names) != arr.shape[1]:
        raise ValueError(f"`colnames` has {len(names)} entries but `{arg}` has {arr.shape[1]} columns")
    if not (np.issubdtype(arr.dtype, np.floating) or np.issubdtype(arr.dtype, np.integer)):
        raise TypeError(f"`{arg}` must be a numeric type.")
    return np.asarray(arr, dtype=np.float64), names


def setup_missing_matrix(data_matrix, global_na):
    """R/utils.R:1-23.  Used for `keep`, `n_good` and the odd corner where global_na has no NA."""
    data = np.asarray(data_matrix, dtype=np.float64)
    excl = np.zeros(data.shape, dtype=bool)
    g = [float(v) for v in global_na]
    if any(np.isnan(v) for v in g):
        excl |= np.isnan(data)
    if any(np.isinf(v) for v in g):
        excl |= np.isinf(data)
    for v in g:
        if np.isfinite(v):
            excl |= (data == v)
    return excl


def setup_comparisons(samples, include_only=None, diag_good=True, include_arg="include_only"):
    """R/kendalltau.R:181-247.  Returns (pi, pj, all_pairs) with 0-based indices; all_pairs is
    True when the list is exactly combn order (+ diagonal), so no explicit list is needed."""
    n_sample = len(samples)
    iu = np.triu_indices(n_sample, k=1)
    pi, pj = iu[0].astype(np.int32), iu[1].astype(np.int32)
    if not diag_good:
        d = np.arange(n_sample, dtype=np.int32)
        pi, pj = np.concatenate([pi, d]), np.concatenate([pj, d])
    all_pairs = True
    if include_only is not None:
        all_pairs = False
        index = {}
        for k, s in enumerate(samples):
            index.setdefault(s, k)
        try:
            import pandas as pd
            if isinstance(include_only, pd.DataFrame):
                include_only = [list(include_only.iloc[:, c]) for c in range(include_only.shape[1])]
        except ImportError:
            pass
        if isinstance(include_only, dict):
            include_only = list(include_only.values())
        # a list mixing scalars and vectors, the documented list(g1 = "s1", g2 = c("s2", "s3")) form
        # (R/kendalltau.R:86-91): scalars are vectors of length one and paste0 recycles them
        if isinstance(include_only, (list, tuple)) and \
                any(isinstance(v, (list, tuple, np.ndarray)) for v in include_only):
            include_only = [list(v) if isinstance(v, (list, tuple, np.ndarray)) else [v] for v in include_only]
        is_list_of_vectors = isinstance(include_only, (list, tuple)) and len(include_only) > 0 and \
            all(isinstance(v, (list, tuple, np.ndarray)) for v in include_only)
        if is_list_of_vectors:
            if len(include_only) != 2:
                raise ValueError(f"`{include_arg}` must be a vector, a data.frame with two columns, "
                                 f"or list of two vectors. Currently, `length({include_arg})` returns "
                                 f"{len(include_only)}")
            l1, l2 = list(include_only[0]), list(include_only[1])
            m = max(len(l1), len(l2))  # paste0 recycles the shorter vector
            l1 = [l1[k % len(l1)] for k in range(m)]
            l2 = [l2[k % len(l2)] for k in range(m)]
            want = set()
            for a, b in zip(l1, l2):
                if a in index and b in index:
                    want.add((index[a], index[b]))
                    want.add((index[b], index[a]))
            code = pi.astype(np.int64) * n_sample + pj
            wcode = np.array([a * n_sample + b for a, b in want], dtype=np.int64)
            keep = np.isin(code, wcode)
        else:
            inc = include_only if isinstance(include_only, (list, tuple, np.ndarray)) else [include_only]
            idx = np.array([index[s] for s in inc if s in index], dtype=np.int32)
            keep = np.isin(pi, idx) | np.isin(pj, idx)
        pi, pj = pi[keep], pj[keep]
    if pi.size == 0:
        raise ValueError("No comparisons to do. Check the list of column names in "
                         f"`{include_arg}` vs those in the samples.")
    return pi, pj, all_pairs


def ici_kt(x, y, perspective="local", alternative="two.sided", continuity=False, output="simple",
           device=0):
    """Information-content-informed Kendall-tau of two vectors (missing = NaN).

    Returns the named vector tau, pvalue, tau_max, completeness.  Degenerate inputs return
    NaN in all four with the reference's warning; a length mismatch raises.
    """
    x = np.asarray(x, dtype=np.float64).ravel()
    y = np.asarray(y, dtype=np.float64).ravel()
    if x.size != y.size:
        raise ValueError("'X' and 'Y' are not the same length!")  # src/kendallc.cpp:168-170
    if x.size == 0:
        return IciKtResult(zip(IciKtResult.NAMES, [np.nan] * 4))  # all-NA branch, :193-199
    data = np.column_stack([x, y])
    r = _lib.run_pairs(data, (), pi=[0], pj=[1], perspective=perspective, alternative=alternative,
                       continuity=continuity, device=device)
    _warn_status(r["status"])
    res = IciKtResult(tau=r["raw"][0], pvalue=r["pvalue"][0], tau_max=r["taumax"][0],
                      completeness=r["completeness"][0])
    if output != "simple":
        print("\n".join(f"{k}: {v}" for k, v in res.items()))
    return res


def _reshape(names, pi, pj, cols):
    C = len(names)
    out = {}
    for k, v in cols.items():
        m = np.zeros((C, C))
        m[pi, pj] = v
        m[pj, pi] = v
        out[k] = m
    return out


def ici_kendalltau(data_matrix, global_na=(np.nan, np.inf, 0), perspective="global", scale_max=True,
                   diag_good=True, include_only=None, alternative="two.sided", continuity=False,
                   check_timing=False, return_matrix=True, colnames=None, device=0, n_gpus=1):
    """All-pairs ICI-Kendall-tau between the columns (samples) of a features x samples matrix.

    Returns a dict with `cor`, `raw`, `pvalue`, `taumax`, `completeness` (C x C matrices),
    `keep`, `run_time` and `names`; with return_matrix=False a long table `cor` (dict of columns
    s1, s2, raw, pvalue, taumax, completeness, cor) like the reference's data.frame.
    n_gpus > 1 spreads the pair order over that many GPUs of the box (the role of the reference's
    furrr workers), still in one library call.
    """
    data, names = _colnames_of(data_matrix, colnames, "data_matrix")
    n, C = data.shape
    log.info("Processing missing values ...")
    global_na = [float(v) for v in global_na]
    exclude_loc = setup_missing_matrix(data, global_na)
    log.info("Figuring out comparisons to do ...")
    if include_only is None and return_matrix and not check_timing:
        # every pair, in the library's own order, straight into matrices: no pair list on the host
        pi = pj = None
        all_pairs = True
        if C < 2 and diag_good:
            raise ValueError("No comparisons to do. Check the list of column names in "
                             "`include_only` vs those in the samples.")
    else:
        pi, pj, all_pairs = setup_comparisons(names, include_only, diag_good)
        n_todo = pi.size

    kw = dict(perspective=perspective, alternative=alternative, continuity=continuity, device=device)
    if check_timing:  # R/kendalltau.R:141-148, 633-669
        rng = np.random.default_rng()
        sel = rng.choice(n_todo, size=min(5, n_todo), replace=False)
        t0 = time.perf_counter()
        _lib.run_pairs(data, global_na, pi=pi[sel], pj=pj[sel], **kw)
        t_total = time.perf_counter() - t0
        t_each = t_total / sel.size
        t_all = t_each * n_todo
        return dict(which=["n_tested", "n_todo", "time_tested", "time_single", "time_all",
                           "time_across_cores", "time_minutes", "time_hours", "time_days"],
                    value=[sel.size, n_todo, t_total, t_each, t_all, t_all, t_all / 60,
                           t_all / 3600, t_all / 216000])

    log.info("Running correlations ...")
    t1 = time.perf_counter()
    if return_matrix and (n_gpus <= 1 or all_pairs):
        # scale_and_reshape on the device: the five C x C matrices come back filled (:357-421)
        n_good = (~exclude_loc).sum(axis=0)
        if all_pairs and n_gpus > 1:  # every GPU fills and returns its own block of columns
            r = _lib.run_matrices(data, global_na, scale_max, diag_good, n_good, devices=range(n_gpus), **kw)
        elif all_pairs:
            r = _lib.run_matrices(data, global_na, scale_max, diag_good, n_good, **kw)
        else:
            r = _lib.run_matrices(data, global_na, scale_max, diag_good, n_good, pi=pi, pj=pj, **kw)
        run_time = time.perf_counter() - t1
        _warn_status_counts(r["status_counts"])
        log.info("Generating the output matrix ...")
        out = {k: r[k] for k in _lib.MATRIX_NAMES}
        out["keep"] = (~exclude_loc).T
        out["run_time"] = run_time
        out["names"] = names
        return out
    if all_pairs and n_gpus > 1:  # one call, the pair order sliced over the GPUs inside the library
        r = _lib.run_pairs(data, global_na, include_diag=not diag_good, devices=range(n_gpus), **kw)
    elif all_pairs:
        r = _lib.run_pairs(data, global_na, include_diag=not diag_good, **kw)
    else:
        r = _lib.run_pairs(data, global_na, pi=pi, pj=pj, **kw)
    run_time = time.perf_counter() - t1
    _warn_status(r["status"])

    log.info("Recombining results ...")
    raw, pvalue, taumax, completeness = r["raw"], r["pvalue"], r["taumax"], r["completeness"]
    if scale_max:  # R/kendalltau.R:368-372: max over the pairs actually computed, na.rm
        cor = raw / r["max_taumax"]
    else:
        cor = raw.copy()
    n_good = (~exclude_loc).sum(axis=0)
    if diag_good:  # R/kendalltau.R:374-386 (appended after scaling, never scaled)
        d = np.arange(C, dtype=np.int32)
        dg = n_good / n_good.max()
        pi, pj = np.concatenate([pi, d]), np.concatenate([pj, d])
        raw, cor = np.concatenate([raw, dg]), np.concatenate([cor, dg])
        pvalue = np.concatenate([pvalue, np.zeros(C)])
        taumax = np.concatenate([taumax, np.ones(C)])
        completeness = np.concatenate([completeness, n_good / n])
    if not return_matrix:
        nm = np.asarray(names, dtype=object)
        return dict(cor=dict(s1=nm[pi], s2=nm[pj], raw=raw, pvalue=pvalue, taumax=taumax,
                             completeness=completeness, cor=cor), run_time=run_time, names=names)
    log.info("Generating the output matrix ...")
    out = _reshape(names, pi, pj, dict(cor=cor, raw=raw, pvalue=pvalue, taumax=taumax,
                                        completeness=completeness))
    out["keep"] = (~exclude_loc).T
    out["run_time"] = run_time
    out["names"] = names
    return out


_NA_METHODS = ("all.obs", "complete.obs", "pairwise.complete.obs", "everything", "na.or.complete")


def _match_arg(use):
    hits = [m for m in _NA_METHODS if m.startswith(use)]
    if use in _NA_METHODS:
        return use
    if len(hits) != 1:
        raise ValueError("'arg' should be one of " + ", ".join(f"'{m}'" for m in _NA_METHODS))
    return hits[0]


def kt_fast(x, y=None, use="everything", alternative="two.sided", continuity=False,
            return_matrix=True, colnames=None, device=0):
    """Plain Kendall-tau-b through the ICI kernel (R/kendalltau.R:448-545).

    As in the reference, `alternative` and `continuity` are accepted but not forwarded to the
    pair kernel (kt_split calls ici_kt(tmp_x, tmp_y) with defaults, :341), and the (i,i) pairs
    are computed (diag_good = FALSE, :479).
    """
    na_method = _match_arg(use)
    if na_method == "na.or.complete":
        raise ValueError("'na.or.complete' is not a supported value for `use`. Please use one of "
                         "all.obs complete.obs pairwise.complete everthing.")
    xa = np.asarray(x) if not hasattr(x, "columns") else None
    if y is None:
        if xa is not None and xa.ndim < 2:
            raise ValueError("`x` and `y` should both be provided as vectors, or `x` should be "
                             "matrix-like. `x` is a single vector, and `y` is `NULL`.")
        data, names = _colnames_of(x, colnames, "x")
    else:
        ya = np.asarray(y)
        if (xa is None or xa.ndim > 1) or ya.ndim > 1:
            raise ValueError("Both `x` and `y` must be vectors.")
        data = np.column_stack([np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)])
        names = list(colnames) if colnames is not None else ["x", "y"]
    n, C = data.shape
    na_vals = np.isnan(data)
    any_na = bool(na_vals.any())
    no_na_rows = na_vals.sum(axis=1) == 0
    pi, pj, _ = setup_comparisons(names, None, diag_good=False)
    P = pi.size
    do = True
    if na_method in ("everything", "all.obs") and any_na:
        do = False
    if na_method == "complete.obs":  # :491 (the reference misspells the pairwise alternative)
        if no_na_rows.sum() == 0:
            do = False
        else:
            data = data[no_na_rows, :]
    tau = np.full(P, np.nan)
    pv = np.full(P, np.nan)
    t1 = time.perf_counter()
    if do:
        kw = dict(perspective="local", alternative="two.sided", continuity=False, device=device)
        if na_method == "pairwise.complete.obs" and np.isnan(data).any():
            # Rows with a missing value in either column are dropped per pair (:323-331).  The pair
            # kernel does that on the device (complete-observations mode); only pairs it flags as
            # unsupported (a column whose missing rows tie with its minimum in fp64) are filtered
            # here and run as two-column problems.
            r = _lib.run_pairs(data, (), include_diag=True, **dict(kw, perspective="complete"))
            tau, pv = r["raw"].copy(), r["pvalue"].copy()
            redo = np.nonzero(r["status"] == 9)[0]
            r["status"][redo] = 0
            _warn_status(r["status"])
            for k in redo:
                good = ~np.isnan(data[:, pi[k]]) & ~np.isnan(data[:, pj[k]])
                tau[k] = pv[k] = np.nan
                if good.sum() == 0:
                    continue
                sub = np.column_stack([data[good, pi[k]], data[good, pj[k]]])
                r2 = _lib.run_pairs(sub, (), pi=[0], pj=[1], **kw)
                _warn_status(r2["status"])
                tau[k], pv[k] = r2["raw"][0], r2["pvalue"][0]
        else:
            r = _lib.run_pairs(data, (), include_diag=True, **kw)
            _warn_status(r["status"])
            tau, pv = r["raw"], r["pvalue"]
    run_time = time.perf_counter() - t1 if do else 0.0
    if not return_matrix:
        nm = np.asarray(names, dtype=object)
        return dict(tau=dict(s1=nm[pi], s2=nm[pj], tau=tau, pvalue=pv), run_time=run_time, names=names)
    out = _reshape(names, pi, pj, dict(tau=tau, pvalue=pv))
    out["run_time"] = run_time
    out["names"] = names
    return out


def pairwise_completeness(data_matrix, global_na=(np.nan, np.inf, 0), include_only=None,
                          return_matrix=True, colnames=None, device=0):
    """1 - (rows missing in either sample) / n for every pair incl. (i,i) (R/kendalltau.R:563-629).

    A by-product of the pair kernel's missing-row masks: the global-perspective completeness of
    the ICI path is exactly this quantity (tests/testthat/test-kendall-tau.R:138-151).
    """
    data, names = _colnames_of(data_matrix, colnames, "data_matrix")
    n, C = data.shape
    global_na = [float(v) for v in global_na]
    # missing-row bit masks and popc(x | y) on the device; no pair kernel runs
    if include_only is None and return_matrix:  # the pair order itself is not needed
        return _lib.pairwise_completeness(data, global_na, want_matrix=True, want_pairs=False, device=device)["matrix"]
    pi, pj, all_pairs = setup_comparisons(names, include_only, diag_good=False)
    if all_pairs:
        r = _lib.pairwise_completeness(data, global_na, device=device)
    else:
        r = _lib.pairwise_completeness(data, global_na, pi=pi, pj=pj, device=device)
    missing, comp = r["missing"].astype(np.float64), r["completeness"]
    if not return_matrix:
        nm = np.asarray(names, dtype=object)
        return dict(s1=nm[pi], s2=nm[pj], missingness=missing, completeness=comp)
    return _reshape(names, pi, pj, dict(completeness=comp))["completeness"]
