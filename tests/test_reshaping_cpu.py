"""Result formats: the cases of the reference's tests/testthat/test-reshaping.R."""
import itertools

import numpy as np
import pytest

from icikendalltau_b200.reshaping import cor_matrix_2_long_df, long_df_2_cor_matrix


def test_round_trip_and_half_tables():
    n = 200
    names = [f"s{i + 1}" for i in range(n)]
    m = np.random.default_rng(1).normal(size=(n, n))
    df = cor_matrix_2_long_df(m, names, names)
    assert df["cor"].size == n * n
    k = [i for i in range(n * n) if df["s1"][i] == "s2" and df["s2"][i] == "s142"]
    assert len(k) == 1 and df["cor"][k[0]] == m[1, 141]

    back, rows, cols = long_df_2_cor_matrix(df)
    assert rows == cols == sorted(names)
    assert back[rows.index("s2"), cols.index("s142")] == m[1, 141]
    assert not np.isnan(back).any()

    # one triangle only: mirrored when square
    pairs = list(itertools.combinations(range(n), 2))
    short = dict(s1=[names[i] for i, _ in pairs], s2=[names[j] for _, j in pairs],
                 cor=np.array([m[i, j] for i, j in pairs]))
    sq, rows, cols = long_df_2_cor_matrix(short)
    assert sq[rows.index("s2"), cols.index("s142")] == m[1, 141]
    assert sq[rows.index("s142"), cols.index("s2")] == m[1, 141]
    ns, rows, cols = long_df_2_cor_matrix(short, is_square=False)
    assert ns.shape == (n - 1, n - 1)
    assert ns[rows.index("s2"), cols.index("s142")] == m[1, 141]

    bad = dict(s1=short["s1"], s2=short["s2"], raw=short["cor"])
    with pytest.raises(ValueError, match="must contain the names"):
        long_df_2_cor_matrix(bad)
    bad["cor"] = bad["raw"]
    again, _, _ = long_df_2_cor_matrix(bad)
    np.testing.assert_array_equal(again, sq)


def test_long_table_of_ici_kendalltau_shape():
    """The long table ici_kendalltau(return_matrix=False) returns feeds long_df_2_cor_matrix."""
    tbl = dict(s1=["a", "a", "b", "a", "b", "c"], s2=["b", "c", "c", "a", "b", "c"],
               cor=[0.5, 0.25, 0.125, 1.0, 1.0, 1.0])
    m, rows, cols = long_df_2_cor_matrix(tbl)
    assert rows == ["a", "b", "c"]
    np.testing.assert_array_equal(m, np.array([[1, 0.5, 0.25], [0.5, 1, 0.125], [0.25, 0.125, 1]]))
