"""Chunked GPU index builder (pqt_assign_bins / pqt_set_db_from_bins / pqt_line_dist_*) against
the oracle's builder: bit-identical bins, lists and line codes, for float and uint8 rows, any
chunking, and for bin-range shards that each encode only their own vectors
(test/test1B.cpp:783-871 is the reference's chunk-wise accumulate)."""
import numpy as np
import pytest

import conftest
import pqt_oracle as po
from util import oracle_query

pytestmark = pytest.mark.gpu


def _handle(c, **params):
    import pqt_b200
    prm = c["prm"]
    t = pqt_b200.PerturbationProTree(prm.dim, prm.p)
    t.set_params(hash_size=prm.hash_size, k1_build=min(16, prm.c1), **params)
    t.setTree(c["cb1"], c["cb2"].reshape(prm.p, prm.c1, prm.c2, -1))
    return t


def _chunked_build(t, X, LP, chunk, on_device, want_lines=True):
    import torch
    N = X.shape[0]
    bins = np.zeros(N, np.uint32)
    for i0 in range(0, N, chunk):
        n = min(chunk, N - i0)
        rows = X[i0:i0 + n]
        if on_device:
            rows = torch.from_numpy(np.ascontiguousarray(rows)).cuda()
        bins[i0:i0 + n] = t.assignBins(rows, n)
    t.setDBFromBins(bins, N)
    t.lineDistBegin(N, LP)
    lines = np.zeros((N, LP), np.uint32) if want_lines else None
    for i0 in range(0, N, chunk):
        n = min(chunk, N - i0)
        rows = X[i0:i0 + n]
        if on_device:
            rows = torch.from_numpy(np.ascontiguousarray(rows)).cuda()
        t.lineDistChunk(rows, i0, n, lines[i0:i0 + n] if want_lines else None)
    t.lineDistEnd()
    return bins, lines


@pytest.fixture(scope="module")
def case_c32():
    # the SIFT-shaped tree: c1 = c2 = 32 (fast kernels with compile-time strides), LP = 32
    return conftest.make_case(N=9000, QN=32, c1=32, c2=32, LP=32, hash_size=200003, seed=5)


@pytest.mark.parametrize("dtype,on_device,chunk", [("u8", True, 2048), ("f32", True, 5000),
                                                   ("u8", False, 3333), ("f32", False, 100000)])
def test_chunked_build_matches_oracle_builder(case_small, dtype, on_device, chunk):
    c = case_small
    prm = c["prm"]
    X = c["X"].astype(np.uint8) if dtype == "u8" else c["X"]
    assert np.array_equal(X.astype(np.float32), c["X"])  # synthetic data is integer valued
    t = _handle(c)
    bins, lines = _chunked_build(t, X, prm.line_parts, chunk, on_device)
    assert np.array_equal(bins, c["bin_of"])
    prefix, counts, db_idx = t.getDB()
    assert np.array_equal(counts, c["counts"])
    assert np.array_equal(prefix, c["prefix"])
    assert np.array_equal(db_idx, c["db_idx"])
    assert np.array_equal(lines, c["lines"])
    assert np.array_equal(t.getLine(), c["lines"])
    # the resident layout: row r of the bin-ordered list holds vector dbIdx[r]
    assert np.array_equal(t.getCodesBinOrder(), c["lines"][c["db_idx"]])
    d0, i0 = oracle_query(c, 256)
    i1, d1 = t.queryKNN(c["Q"], c["Q"].shape[0], 256)
    assert np.array_equal(i1, i0) and np.array_equal(d1, d0)
    t.close()


@pytest.mark.parametrize("which", ["c32_lp32", "lp32", "lp8"])
def test_fast_kernel_shapes(which, case_c32, case_lp32):
    c = {"c32_lp32": case_c32, "lp32": case_lp32}.get(which) or \
        conftest.make_case(N=5000, QN=8, LP=8, hash_size=65537, seed=9)
    prm = c["prm"]
    t = _handle(c)
    bins, lines = _chunked_build(t, c["X"].astype(np.uint8), prm.line_parts, 4096, True)
    assert np.array_equal(bins, c["bin_of"])
    assert np.array_equal(lines, c["lines"])
    prefix, counts, db_idx = t.getDB()
    assert np.array_equal(db_idx, c["db_idx"])
    t.close()


def test_generic_shapes_fall_back(case_small):
    """c1 = 8 (below the warp kernels' shapes) takes the CTA-per-vector kernels"""
    c = conftest.make_case(N=3000, QN=8, c1=8, c2=8, LP=16, hash_size=65537, seed=11)
    t = _handle(c)
    bins, lines = _chunked_build(t, c["X"].astype(np.uint8), 16, 1000, True)
    assert np.array_equal(bins, c["bin_of"])
    assert np.array_equal(lines, c["lines"])
    t.close()


def test_unusual_values_take_the_ieee_division(case_small):
    """segments that coincide with a centroid (distance 0) and tiny / huge coordinates leave
    the range of the hoisted-reciprocal quotient: the kernel must fall back per vector"""
    c = case_small
    prm = c["prm"]
    X = c["X"][:512].copy()
    X[0, :8] = c["cb1"][3, :8]          # zero distance to centroid 3 on line part 0
    X[1] *= 1e-7
    X[2] *= 3e5
    X[3, 8:16] = c["cb1"][5, 8:16] + 1e-6
    idx = po.build_index(prm, c["cb1"], c["cb2"], X, k1_build=min(16, prm.c1))
    t = _handle(c)
    bins, lines = _chunked_build(t, X, prm.line_parts, 200, True)
    assert np.array_equal(bins, idx["bin_of"])
    assert np.array_equal(lines, idx["lines"])
    t.close()


def test_sharded_build_keeps_only_the_own_slice(case_small):
    """every shard runs both passes over all chunks but encodes and stores only the vectors of
    its own bin-range slice; the slices concatenate to the unsharded code array"""
    import pqt_b200  # noqa: F401
    c = case_small
    prm = c["prm"]
    N = c["X"].shape[0]
    X8 = c["X"].astype(np.uint8)
    full = c["lines"][c["db_idx"]]
    world = 3
    for rank in range(world):
        t = _handle(c)
        t.setShard(rank, world)
        _chunked_build(t, X8, prm.line_parts, 7000, True, want_lines=False)
        lo, hi = (N * rank) // world, (N * (rank + 1)) // world
        got = np.zeros((hi - lo, prm.line_parts), np.uint32)
        t.getCodesBinOrder(0, hi - lo, got)
        assert np.array_equal(got, full[lo:hi])
        t.close()


def test_build_errors(case_small):
    import pqt_b200
    c = case_small
    prm = c["prm"]
    t = _handle(c)
    N = c["X"].shape[0]
    with pytest.raises(pqt_b200.PqtError):
        t.lineDistBegin(N, prm.line_parts)  # no DB yet
    bad = c["bin_of"].copy()
    bad[7] = prm.hash_size
    with pytest.raises(pqt_b200.PqtError):
        t.setDBFromBins(bad, N)
    t.setDBFromBins(c["bin_of"], N)
    with pytest.raises(pqt_b200.PqtError):
        t.lineDistChunk(c["X"][:10], 0, 10)  # begin missing
    t.lineDistBegin(N, prm.line_parts)
    with pytest.raises(pqt_b200.PqtError):
        t.lineDistChunk(c["X"][:10], N - 5, 10)  # past the end
    with pytest.raises(pqt_b200.PqtError):
        t.queryKNN(c["Q"], 4, 16)  # codes not complete
    t.close()
