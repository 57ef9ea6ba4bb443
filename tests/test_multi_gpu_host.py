"""N>1 host logic on CPU: two gloo processes play two bin-range shards.  Every rank dispatches
the candidates of its own queries to the shard that holds them (the partition rule of
dispatch_kernel, restated in pqt_b200.sharding.dispatch), every shard evaluates its inbox with
the oracle's ADC and returns the distances to the query's owner, and the owners rank with the
reference network: the result must be the single-index queryKNN result exactly."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import conftest  # noqa: F401
import pqt_oracle as po
from pqt_b200 import sharding
from util import oracle_query


def _candidates(case, k):
    """candidate positions (bin order) and counts of Step E1, plus the oracle's answer"""
    d, i, st = oracle_query(case, k, stages=True)
    QN = case["Q"].shape[0]
    mv = po.pow2ceil(k)
    inv = np.empty(case["db_idx"].size, np.int64)
    inv[case["db_idx"]] = np.arange(case["db_idx"].size)
    pos = np.zeros((QN, mv), np.int64)
    for q in range(QN):
        nv = int(st["n_vec"][q])
        pos[q, :nv] = inv[st["select_idx"][q, :nv]]
    return pos, st["n_vec"].astype(np.int64), st["lut"], d, i


def _worker(rank, world, port, case, pos, nvec, lut, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prm = case["prm"]
    n_db = case["db_idx"].size
    QN, mv = pos.shape
    qlo, qhi = sharding.query_slice(QN, rank, world)
    lo, hi = sharding.shard_bounds(n_db, rank, world)
    # 1. dispatch the own queries' candidates: one inbox row per (shard, query)
    rows = sharding.dispatch(torch.from_numpy(pos[qlo:qhi]), torch.from_numpy(nvec[qlo:qhi]), n_db, world)
    send = [[(r[0].numpy(), r[1].numpy()) for r in rows[dst]] for dst in range(world)]
    inbox = [None] * world
    # (gloo has no all-to-all for objects: gather everything, keep the own part)
    gathered = [None] * world
    dist.all_gather_object(gathered, send)
    for src in range(world):
        inbox[src] = gathered[src][rank]
    # 2. scan the own inbox: codes of the own slice only, LUT of the query recomputed locally
    codes_local = case["lines"][case["db_idx"][lo:hi]]  # bin-ordered slice
    results = []
    for src in range(world):
        s_lo, _ = sharding.query_slice(QN, src, world)
        per_q = []
        for ql, (lpos, ent) in enumerate(inbox[src]):
            assert np.all((lpos >= 0) & (lpos < hi - lo))
            q = s_lo + ql
            vals = np.array([po.line_adc(prm, lut[q], case["cb_dist"], codes_local[p]) for p in lpos],
                            np.float32)
            per_q.append((ent, vals))
        results.append(per_q)
    # 3. distances back to the owners
    back = [None] * world
    dist.all_gather_object(back, results)
    # 4. rank the own queries
    res = []
    for ql in range(qhi - qlo):
        q = qlo + ql
        val = np.full(mv, np.float32(1e7), np.float32)
        idx = np.full(mv, po.PAD_IDX, np.uint32)
        seen = np.zeros(mv, bool)
        for shard in range(world):
            ent, vals = back[shard][rank][ql]
            assert not seen[ent].any()  # every candidate has exactly one evaluator
            seen[ent] = True
            val[ent] = vals
        nv = int(nvec[q])
        assert seen[:nv].all() and not seen[nv:].any()
        idx[:nv] = case["db_idx"][pos[q, :nv]]
        v, ii = po.bitonic(val, idx)
        res.append((v[:k].copy(), ii[:k].copy()))
    out[rank] = (qlo, qhi, res)
    dist.destroy_process_group()


def test_two_shards_return_the_single_gpu_result(case_small):
    c = dict(case_small)
    c["Q"] = case_small["Q"][:16]
    k = 128
    pos, nvec, lut, d0, i0 = _candidates(c, k)
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    slim = {kk: c[kk] for kk in ("prm", "db_idx", "lines", "cb_dist")}
    mp.spawn(_worker, args=(world, port, slim, pos, nvec, lut, k, out), nprocs=world, join=True)
    for r in range(world):
        qlo, qhi, res = out[r]
        for q in range(qlo, qhi):
            v, ii = res[q - qlo]
            assert np.array_equal(v, d0[q]) and np.array_equal(ii, i0[q])


def test_shard_bounds_cover_the_list_without_overlap():
    for n in (1, 7, 1000000, 999999937):
        for world in (1, 2, 3, 8):
            edges = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and a <= b
    assert sharding.query_slice(10000, 3, 8) == (3750, 5000)
    with pytest.raises(AssertionError):
        sharding.query_slice(10, 0, 3)


def test_dispatch_rule_matches_the_shard_bounds():
    """dispatch_kernel finds the shard of a candidate position by counting lower bounds; that
    must agree with the slices pqt_set_shard keeps, for every position and world size."""
    for n in (1, 7, 1000, 12345, 1000003):
        for world in range(1, 9):
            bounds = [sharding.shard_bounds(n, r, world) for r in range(world)]
            probe = sorted(set([0, n - 1] + [b for lo, hi in bounds for b in (lo - 1, lo, hi - 1, hi)
                                             if 0 <= b < n]))
            for pos in probe:
                r = sharding.shard_of(pos, n, world)
                lo, hi = bounds[r]
                assert lo <= pos < hi, (n, world, pos, r)
    # the torch reference of the dispatch partitions every list exactly
    lp = torch.tensor([[5, 0, 99, 42, 42, 7, 0, 0]], dtype=torch.int64)
    rows = sharding.dispatch(lp, torch.tensor([6]), 100, 4)
    got = sorted((int(e), r) for r in range(4) for e in rows[r][0][1])
    assert [e for e, _ in got] == list(range(6))
    for r in range(4):
        lo, hi = sharding.shard_bounds(100, r, 4)
        assert all(0 <= int(p) < hi - lo for p in rows[r][0][0])
