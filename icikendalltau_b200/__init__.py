"""icikendalltau_b200 -- B200-native (sm_100a) all-pairs ICI-Kendall-tau.

Host-side mirror of the reference's R interface (`ici_kendalltau`, `ici_kt`, `kt_fast`,
`pairwise_completeness`) over the C ABI of libicikt_b200.so.  No CPU fallback.
"""
from ._lib import IciktError, Plan, load, pnorm_device, run_pairs  # noqa: F401
from .reshaping import cor_matrix_2_long_df, long_df_2_cor_matrix  # noqa: F401
from .api import (IciKtResult, ici_kendalltau, ici_kt, kt_fast, pairwise_completeness,  # noqa: F401
                  setup_comparisons, setup_missing_matrix)

__all__ = ["cor_matrix_2_long_df", "long_df_2_cor_matrix", "ici_kendalltau", "ici_kt", "kt_fast", "pairwise_completeness", "run_pairs", "Plan",
           "IciktError", "IciKtResult", "load", "pnorm_device", "setup_comparisons",
           "setup_missing_matrix"]
