"""The pair kernel's launch shapes (icikt_launch_shape, host only): every supported length gets a shape that fits
an SM, the in-place variant for long vectors always has a counter / staging area behind its sequence buffer, and
the global scratch is only what is left."""
import numpy as np
import pytest

from icikendalltau_b200 import _lib

SMEM_MAX = 227 * 1024
LENGTHS = sorted(set([1, 2, 31, 33, 255, 257, 1000, 2000, 2049, 5000, 6887, 8192, 8193, 16384, 20000, 22528, 22529,
                      28672, 30000, 32769, 40000, 45000, 50000, 52000, 57344, 60000, 64511, 64512, 64513, 65535] +
                     list(range(3000, 65535, 3701))))


def fixed_smem(n, warps, runs):
    """tiled_smem_bytes(0, wstride, fmask_words) of icikt_pairs.cu"""
    nwords = ((n + 31) & ~31) // 32
    wstride = (nwords + 3) & ~3
    fwords = (warps * runs * 8 + 3) & ~3
    return 8 * 32 * 4 + 16 + 256 + 4 * wstride + 128 + 4 * fwords


@pytest.mark.parametrize("tier", [0, 1, 2])
def test_every_length_has_a_shape_that_fits(tier):
    for n in LENGTHS:
        s = _lib.launch_shape(n, tier)
        assert 1 <= s["warps"] <= 32 and s["runs"] >= 1
        assert s["cap"] == s["warps"] * s["runs"] * 256 and n <= s["cap"] <= 65536, (n, s)  # 16-bit slot indices
        if s["variant"] != "gmem":
            assert fixed_smem(n, s["warps"], s["runs"]) + s["region_bytes"] <= SMEM_MAX, (n, s)
        if s["variant"] == "smem":
            assert s["region_bytes"] >= (4, 5, 8)[tier] * s["cap"], (n, s)  # two u16 buffers (+ counters / pass B)
        if tier == 2:
            assert s["variant"] != "inplace"  # pass B has no in-place form


def test_inplace_variant_always_has_its_staging_area():
    """In place: one u16 sequence buffer; everything else of the 227 KB is the counter area of the tie groups,
    which first stages the other column's rank table for the gather in at most three parts."""
    seen = 0
    for tier in (0, 1):
        for n in LENGTHS:
            s = _lib.launch_shape(n, tier)
            if s["variant"] != "inplace":
                continue
            seen += 1
            assert s["warps"] <= 28 and s["runs"] == 9
            area = s["region_bytes"] - 2 * s["cap"]
            assert area >= s["cap"], (n, s)                       # at least the round-1 counter area
            assert s["stage_rows"] > 0 and s["stage_rows"] % 64 == 0 and 2 * s["stage_rows"] <= area, (n, s)
            assert -(-n // s["stage_rows"]) <= 3, (n, s)
            assert fixed_smem(n, s["warps"], s["runs"]) + s["region_bytes"] > SMEM_MAX - 64  # all of the SM is used
    assert seen >= 8
    assert _lib.launch_shape(60000, 1)["variant"] == "inplace"      # config 4
    assert _lib.launch_shape(20000, 0)["variant"] == "smem"         # the target
    assert _lib.launch_shape(20000, 0)["warps"] == 16 and _lib.launch_shape(20000, 0)["runs"] == 5
    assert _lib.launch_shape(65535, 0)["variant"] == "gmem"
    assert _lib.launch_shape(60000, 0, complete_obs=True)["variant"] == "gmem"  # that mode has no in-place kernel


def test_launch_shape_rejects_bad_arguments():
    with pytest.raises(_lib.IciktError):
        _lib.launch_shape(0)
    with pytest.raises(_lib.IciktError):
        _lib.launch_shape(70000)
    with pytest.raises(_lib.IciktError):
        _lib.launch_shape(100, tier=3)
