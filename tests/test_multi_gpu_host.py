"""N>1 host logic on CPU: two gloo processes emulate two bin-range shards with the
oracle's candidate arrays and must assemble the single-GPU result exactly."""
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


def _candidate_arrays(case, k):
    """Full (unranked) candidate arrays of Step E2 from the oracle's stages."""
    d, i, st = oracle_query(case, k, stages=True)
    prm = case["prm"]
    QN = case["Q"].shape[0]
    mv = po.pow2ceil(k)
    inv = np.empty(case["db_idx"].size, np.int64)
    inv[case["db_idx"]] = np.arange(case["db_idx"].size)
    val = np.full((QN, mv), np.float32(1e7), np.float32)
    idx = np.full((QN, mv), -1, np.int32)  # PAD_IDX as int32
    pos = np.zeros((QN, mv), np.int64)
    for q in range(QN):
        nv = int(st["n_vec"][q])
        ids = st["select_idx"][q, :nv]
        # a bin may be listed twice, but an id has one position in the bin-ordered list
        pos[q, :nv] = inv[ids]
        idx[q, :nv] = ids.astype(np.int32)
        for a in range(nv):
            val[q, a] = po.line_adc(prm, st["lut"][q], case["cb_dist"], case["lines"][ids[a]])
    return val, idx, pos, st["n_vec"].astype(np.int64), d, i


def _worker(rank, world, port, val, idx, pos, nvec, n_db, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_bounds(n_db, rank, world)
    v, i = sharding.mask_to_shard(torch.from_numpy(val), torch.from_numpy(idx),
                                  torch.from_numpy(pos), torch.from_numpy(nvec), lo, hi, rank == 0)
    QN, mv = val.shape
    qlo, qhi = sharding.query_slice(QN, rank, world)
    vo = torch.empty((qhi - qlo, mv), dtype=torch.float32)
    io = torch.empty((qhi - qlo, mv), dtype=torch.int32)
    sharding.exchange(v, i, vo, io, rank, world)
    out[rank] = (vo.numpy().copy(), io.numpy().copy(), qlo, qhi)
    dist.destroy_process_group()


def test_two_shards_assemble_the_single_gpu_candidates(case_small):
    c = dict(case_small)
    c["Q"] = case_small["Q"][:16]
    k = 128
    val, idx, pos, nvec, d0, i0 = _candidate_arrays(c, k)
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, val, idx, pos, nvec, c["db_idx"].size, out), nprocs=world,
             join=True)
    for r in range(world):
        vo, io, qlo, qhi = out[r]
        assert np.array_equal(vo, val[qlo:qhi])
        assert np.array_equal(io, idx[qlo:qhi])
        # ranking the assembled arrays with the reference network gives queryKNN's output
        for q in range(qlo, qhi):
            v, ii = po.bitonic(vo[q - qlo], io[q - qlo].view(np.uint32))
            assert np.array_equal(v[:k], d0[q]) and np.array_equal(ii[:k], i0[q])


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


def test_pull_mode_shard_lookup_matches_the_shard_bounds():
    """The pull-mode kernel finds the shard of a candidate position by counting lower bounds;
    that must agree with the slices pqt_set_shard keeps, for every position and world size."""
    for n in (1, 7, 1000, 12345, 1000003):
        for world in range(1, 9):
            bounds = [sharding.shard_bounds(n, r, world) for r in range(world)]
            probe = sorted(set([0, n - 1] + [b for lo, hi in bounds for b in (lo - 1, lo, hi - 1, hi)
                                             if 0 <= b < n]))
            for pos in probe:
                r = sharding.shard_of(pos, n, world)
                lo, hi = bounds[r]
                assert lo <= pos < hi, (n, world, pos, r)
