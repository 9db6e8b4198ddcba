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


def column_range(C, rank, world):
    """Contiguous slice [lo, hi) of the columns rank `rank` preprocesses (K1 is sharded too)."""
    return C * rank // world, C * (rank + 1) // world


def table_slices(tables, C, world):
    """Byte ranges to exchange: for every table (ptr, bytes_per_column) and rank r the triple
    (table index, byte offset, byte length) of rank r's column slice.  Pure arithmetic (CPU-tested)."""
    out = []
    for t, (_, bpc) in enumerate(tables):
        for r in range(world):
            lo, hi = column_range(C, r, world)
            out.append((t, r, lo * bpc, (hi - lo) * bpc))
    return out


class _DevBytes:
    """A span of device memory as a __cuda_array_interface__ object (torch.as_tensor wraps it)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


def all_gather_columns(full, bpc, C, rank, world, group=None):
    """All-gather of one table between the ranks of a torch.distributed group.  `full` is the whole
    table as a flat uint8 tensor (C columns of `bpc` bytes) in which this rank has filled its own
    column slice; afterwards every rank holds every column.  Equal slices go as one in-place
    all-gather, ragged ones as one broadcast per rank.  Returns the number of collectives."""
    import torch.distributed as dist

    if C % world == 0:
        per = bpc * (C // world)
        dist.all_gather_into_tensor(full, full[rank * per:(rank + 1) * per].clone() if full.device.type == "cpu"
                                    else full[rank * per:(rank + 1) * per], group=group)
        return 1
    calls = 0
    for r in range(world):
        lo, hi = column_range(C, r, world)
        if hi > lo:
            dist.broadcast(full[lo * bpc:hi * bpc], src=r if group is None else dist.get_global_rank(group, r),
                           group=group)
            calls += 1
    return calls


def exchange_tables(plan, rank, world, stream, group=None):
    """All-gather of the K1 tables over NCCL: every rank has filled the slices of its own columns
    (Plan.columns_range); afterwards every rank holds every column.  The collectives are enqueued
    behind the plan's stream (`stream`: a torch ExternalStream of Plan.stream()), so no host
    synchronisation is involved.  Returns the number of collectives issued."""
    import torch

    calls = 0
    dev = torch.device("cuda", torch.cuda.current_device())
    with torch.cuda.stream(stream):
        for ptr, bpc in plan.tables():
            full = torch.as_tensor(_DevBytes(ptr, bpc * plan.C), device=dev)
            calls += all_gather_columns(full, bpc, plan.C, rank, world, group)
    return calls
