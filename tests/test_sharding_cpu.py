"""CPU tests of the multi-GPU host logic (world_size 2 over gloo): pair-range sharding, the
max-over-ranks reduction bench.py uses, and the pair-index helper of the C ABI."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from icikendalltau_b200 import _lib, sharding
from oracle import oracle as O


def test_pair_from_index_matches_combn_order():
    for C in (2, 3, 7, 100, 1001):
        iu = np.triu_indices(C, k=1)
        P = iu[0].size
        idx = np.unique(np.concatenate([np.arange(min(P, 50)), np.arange(max(0, P - 50), P),
                                        np.random.default_rng(C).integers(0, P, size=200)]))
        for k in idx:
            assert _lib.pair_from_index(C, int(k)) == (iu[0][k], iu[1][k])
        assert _lib.pair_from_index(C, P + C - 1, include_diag=True) == (C - 1, C - 1)
    opi, opj = O.setup_comparisons(9, None, diag_good=False)
    for k in range(opi.size):
        assert _lib.pair_from_index(9, k, include_diag=True) == (opi[k], opj[k])


def test_pair_ranges_cover_and_balance():
    for P in (1, 7, 4950, 12497500):
        for world in (1, 2, 3, 4, 8):
            r = [sharding.pair_range(P, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == P
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_column_ranges_cover_and_table_slices():
    for C in (1, 7, 100, 2000, 5001):
        for world in (1, 2, 3, 8):
            r = [sharding.column_range(C, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == C
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    sl = sharding.table_slices([(1000, 40), (5000, 64)], 10, 4)
    assert len(sl) == 8
    for t, bpc in ((0, 40), (1, 64)):
        mine = [(off, ln) for (tt, r, off, ln) in sl if tt == t]
        assert mine[0][0] == 0 and sum(ln for _, ln in mine) == 10 * bpc
        assert all(mine[k][0] + mine[k][1] == mine[k + 1][0] for k in range(3))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank "computes" its slice with the oracle standing in for the GPU (no GPU here);
    # what is under test is the slicing / gathering / max logic around it
    rng = np.random.default_rng(5)
    x = rng.normal(size=(60, 9))
    x[rng.random(x.shape) < 0.2] = np.nan
    pi, pj = O.setup_comparisons(9, None, True)
    P = pi.size
    lo, hi = sharding.pair_range(P, rank, world)
    mine = O.pair_loop(x, pi[lo:hi], pj[lo:hi], perspective="local")
    mine["max_taumax"] = float(np.nanmax(mine["taumax"]))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # the table exchange of the sharded K1 (bench.py runs it over NCCL on device memory): every rank
    # fills its own column slice of a table, afterwards every rank holds every column -- equal
    # slices (one in-place all-gather) and ragged ones (one broadcast per rank)
    tables_ok = True
    for C, bpc in ((8, 24), (9, 10)):
        want = torch.arange(C * bpc, dtype=torch.int64).remainder(251).to(torch.uint8)
        tbl = torch.zeros(C * bpc, dtype=torch.uint8)
        c0, c1 = sharding.column_range(C, rank, world)
        tbl[c0 * bpc:c1 * bpc] = want[c0 * bpc:c1 * bpc]
        calls = sharding.all_gather_columns(tbl, bpc, C, rank, world)
        tables_ok = tables_ok and bool(torch.equal(tbl, want)) and calls == (1 if C % world == 0 else world)
    flags = [None] * world
    dist.all_gather_object(flags, tables_ok)
    if rank == 0:
        full = O.pair_loop(x, pi, pj, perspective="local")
        got = sharding.gather_results(gathered, P)
        ok = all(np.array_equal(got[k], full[k], equal_nan=True) for k in ("raw", "pvalue", "taumax",
                                                                          "completeness", "status"))
        ok = ok and got["max_taumax"] == float(np.nanmax(full["taumax"])) and t.item() == world and all(flags)
        q.put(ok)
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
