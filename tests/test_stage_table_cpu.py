"""The launch table of the pipelined one-shot call (icikt_stage_table, host only): every pair of the combn
order exactly once, a launch never touches a column that has not been uploaded before it, and the slot
ranges declared complete after a launch really are complete and tile the pair order."""
import numpy as np
import pytest

from icikendalltau_b200 import _lib


@pytest.mark.parametrize("C,diag,slots,blocks", [
    (2000, False, 296, 8),   # the north-star target
    (5000, True, 296, 8),    # config 5 with the diagonal pairs
    (130, False, 296, 8),    # barely above the threshold: two chunks, one block
    (600, True, 8, 4),
    (97, False, 4, 8),
    (1000, False, 296, 1),   # no row blocks
    (2, False, 296, 8),
    (1, True, 296, 8),
])
def test_stage_table_covers_the_pair_order(C, diag, slots, blocks):
    U, S = _lib.stage_table(C, diag, slots, blocks)
    ptri = C * (C - 1) // 2
    P = ptri + (C if diag else 0)
    seen = np.zeros(P, dtype=np.int32)
    uploaded = 0       # columns on the device so far
    complete = 0       # slots declared complete so far
    next_unit = 0
    for col_lo, col_hi, unit_lo, unit_hi, slot_lo, slot_hi in S:
        if col_hi > col_lo:
            assert col_lo == uploaded, "column chunks are consecutive"
            uploaded = col_hi
        assert unit_lo == next_unit and unit_hi >= unit_lo, "launches tile the unit table"
        next_unit = unit_hi
        for slot0, i, j0, cnt in U[unit_lo:unit_hi]:
            assert cnt >= 1 and i < uploaded and j0 + cnt - 1 < uploaded, "only columns that have been uploaded"
            if i == j0:
                assert diag and cnt == 1 and slot0 == ptri + i
            else:
                assert j0 > i and j0 + cnt <= C  # a unit stays inside its row of the pair order
                assert slot0 == i * (2 * C - i - 1) // 2 + (j0 - i - 1), "combn order (R/kendalltau.R:213)"
            seen[slot0:slot0 + cnt] += 1
        if slot_hi > slot_lo:
            assert slot_lo == complete, "completed ranges are consecutive"
            complete = slot_hi
            assert (seen[slot_lo:slot_hi] == 1).all(), "a range declared complete is complete"
    assert uploaded == C and complete == P and next_unit == len(U)
    assert (seen == 1).all()


def test_stage_table_unit_lengths_taper():
    """Long units while much work is left, single pairs at the end of every launch (the persistent CTAs
    then run dry together); never more than 16 pairs."""
    U, S = _lib.stage_table(2000, False, 296, 8)
    for _, _, unit_lo, unit_hi, _, _ in S:
        cnt = U[unit_lo:unit_hi, 3]
        assert cnt.max() <= 16
        assert cnt[-1] == 1 or cnt.sum() < 4 * 296
    big = U[S[-1][2]:S[-1][3], 3]
    assert big.max() == 16 and (big[-296:] == 1).all()


def test_stage_table_rejects_bad_arguments():
    with pytest.raises(_lib.IciktError):
        _lib.stage_table(0)
    with pytest.raises(_lib.IciktError):
        _lib.stage_table(10, cta_slots=0)
