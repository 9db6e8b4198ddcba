"""Sharding of the pair order across GPUs: pairs are independent, so every rank
takes one contiguous slice of the combn order -- the reference's own split into
ceil(P/ncore) contiguous chunks (R/kendalltau.R:250-253) -- and writes its own output block.
The only cross-pair quantity is max(taumax) for scale_max, a host-side max of <= 8 scalars."""
from __future__ import annotations

import numpy as np


def n_pairs(C, include_diag=False):
    return C * (C - 1) // 2 + (C if include_diag else 0)


def pair_range(P, rank, world):
    """Contiguous slice [lo, hi) of rank `rank`; slices differ by at most one pair."""
    return P * rank // world, P * (rank + 1) // world


def combine_max_taumax(per_rank_max):
    """max(taumax, na.rm = TRUE) over ranks (R/kendalltau.R:368-370)."""
    v = np.asarray(per_rank_max, dtype=np.float64)
    return float(np.nanmax(v)) if (~np.isnan(v)).any() else float("nan")


def gather_results(per_rank, P):
    """Concatenate per-rank result dicts (each in its slice's order) into the full pair order."""
    out = {}
    for k in ("raw", "pvalue", "taumax", "completeness", "status"):
        out[k] = np.concatenate([r[k] for r in per_rank])
        assert out[k].size == P
    out["max_taumax"] = combine_max_taumax([r["max_taumax"] for r in per_rank])
    return out
