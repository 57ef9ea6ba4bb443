"""BASELINE configs[1] at full size through the C ABI: 1M x 128-d, p=4, c1=c2=32, lineparts=16,
the reference's HASH_SIZE (pqt/PerturbationProTree.hh:12), a 10k-query batch with k=4096 from
HOST buffers (five 2048-query slabs with overlapped result copies), every query compared with
the oracle bit for bit.  The index is built by the GPU builder (itself checked against the
oracle's builder in test_gpu_build.py) and handed to the oracle through getDB / getLine, the way
tool_createdb's files reach tool_query."""
import os

import numpy as np
import pytest

import conftest  # noqa: F401  (sys.path)
import pqt_oracle as po

pytestmark = pytest.mark.gpu

N, QN, K = 1000000, 10000, 4096
DIM, P, C1, C2, LP, HASH = 128, 4, 32, 32, 16, 400000000


@pytest.fixture(scope="module")
def c2_index():
    import pqt_b200
    from pqt_b200 import synth
    mu = synth.centres(4096, DIM)
    X = synth.db_vectors(0, N, DIM, 4096, mu=mu)
    Qu, _ = synth.query_vectors(QN, N, DIM, 4096, mu=mu)
    Q = Qu.astype(np.float32)
    cb1, cb2 = synth.train_tree(X[:20000].astype(np.float32), P, C1, C2, iters=6, seed=77)
    t = pqt_b200.PerturbationProTree(DIM, P)
    t.set_params(hash_size=HASH, k1_build=16)
    t.setTree(cb1, cb2.reshape(P, C1, C2, DIM // P))
    chunk = 250000
    bins = np.zeros(N, np.uint32)
    for i0 in range(0, N, chunk):
        bins[i0:i0 + chunk] = t.assignBins(np.ascontiguousarray(X[i0:i0 + chunk]), min(chunk, N - i0))
    t.setDBFromBins(bins, N)
    t.lineDistBegin(N, LP)
    for i0 in range(0, N, chunk):
        t.lineDistChunk(np.ascontiguousarray(X[i0:i0 + chunk]), i0, min(chunk, N - i0))
    t.lineDistEnd()
    yield dict(t=t, Q=Q, cb1=cb1, cb2=cb2)
    t.close()


def test_c2_full_size_every_query_matches_the_oracle(c2_index):
    t, Q = c2_index["t"], c2_index["Q"]
    idx, dist = t.queryKNN(Q, QN, K)  # numpy in, numpy out: host buffers, slab pipeline
    prefix, counts, db_idx = t.getDB()
    lines = t.getLine()
    prm = po.default_params(DIM, P, C1, C2, LP, hash_size=HASH)
    d0, i0 = po.query_knn(prm, c2_index["cb1"], c2_index["cb2"], prefix, counts, db_idx, lines, Q, K,
                          nthreads=os.cpu_count() or 1)
    assert np.array_equal(idx, i0)
    assert np.array_equal(dist, d0)
    nv = (i0 != po.PAD_IDX).sum(1)
    assert 0 < nv.min() < nv.max() <= K  # ragged candidate lists, pads behind them


def test_c2_full_size_big_variant_sample(c2_index):
    """queryBIGKNNRerank2 on the same index: a 512-query sample against the oracle"""
    t, Q = c2_index["t"], c2_index["Q"][:512]
    idx, dist = t.queryBIGKNNRerank2(Q, 512, K)
    prefix, counts, db_idx = t.getDB()
    lines = t.getLine()
    bp = po.big_params(DIM, P, C1, C2, LP, hash_size=HASH)
    d0, i0, info = po.query_big_knn_rerank2(bp, c2_index["cb1"], c2_index["cb2"], prefix, counts, db_idx,
                                           lines, Q, K, nthreads=os.cpu_count() or 1)
    ok = ~info["ambiguous"]
    assert ok.sum() >= 500
    assert np.array_equal(idx[ok], i0[ok])
    assert np.array_equal(dist[ok], d0[ok])
