"""Result formats either side of the pair path: square matrix <-> long table.

Mirrors `cor_matrix_2_long_df` / `long_df_2_cor_matrix` of the reference (R/reshaping.R:16-68):
the long table is a dict of equally long columns `s1`, `s2`, `cor` (what ici_kendalltau returns
with return_matrix=False); a matrix travels with its row and column names.
"""
from __future__ import annotations

import numpy as np


def cor_matrix_2_long_df(in_matrix, rownames=None, colnames=None):
    """Square (or rectangular) matrix -> long table, column-major like R's `stack()`:
    all rows of the first column first (R/reshaping.R:16-33)."""
    m = np.asarray(in_matrix)
    if m.ndim != 2:
        raise ValueError("`in_matrix` must be a matrix")
    try:  # pandas DataFrame carries its own names
        rownames = list(in_matrix.index) if rownames is None else rownames
        colnames = list(in_matrix.columns) if colnames is None else colnames
    except AttributeError:
        pass
    if rownames is None or colnames is None:
        raise ValueError("row and column names of `in_matrix` must be given")
    rn, cn = np.asarray(list(rownames), dtype=object), np.asarray(list(colnames), dtype=object)
    if rn.size != m.shape[0] or cn.size != m.shape[1]:
        raise ValueError("names do not match the shape of `in_matrix`")
    return dict(s1=np.tile(rn, m.shape[1]), s2=np.repeat(cn, m.shape[0]), cor=m.reshape(-1, order="F").copy())


def long_df_2_cor_matrix(long_df, is_square=True):
    """Long table -> (matrix, rownames, colnames) (R/reshaping.R:45-68).  Names are the sorted
    distinct labels (R's factor levels).  If the table holds fewer rows than the matrix has
    cells (one triangle only) and `is_square`, the mirrored entries are filled too; cells that
    never appear stay NaN."""
    if not all(k in long_df for k in ("s1", "s2", "cor")):
        raise ValueError("The data.frame must contain the names 's1', 's2', and 'cor'.")
    s1 = np.asarray(list(long_df["s1"]), dtype=object)
    s2 = np.asarray(list(long_df["s2"]), dtype=object)
    cor = np.asarray(long_df["cor"], dtype=np.float64)
    if is_square:
        rows = cols = sorted(set(s1.tolist()) | set(s2.tolist()))
    else:
        rows, cols = sorted(set(s1.tolist())), sorted(set(s2.tolist()))
    ri = {k: i for i, k in enumerate(rows)}
    ci = {k: i for i, k in enumerate(cols)}
    out = np.full((len(rows), len(cols)), np.nan)
    i1 = np.fromiter((ri[k] for k in s1), dtype=np.int64, count=s1.size)
    i2 = np.fromiter((ci[k] for k in s2), dtype=np.int64, count=s2.size)
    out[i1, i2] = cor
    if is_square and cor.size != out.size:
        out[np.fromiter((ri[k] for k in s2), dtype=np.int64, count=s2.size),
            np.fromiter((ci[k] for k in s1), dtype=np.int64, count=s1.size)] = cor
    return out, rows, cols
