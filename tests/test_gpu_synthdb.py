"""Bench plumbing on the GPU: the CUDA generator against the numpy generator, the exact
ground truth against the oracle's brute force, tool_synthdb's files against the oracle's
builder."""
import os
import subprocess
import sys

import numpy as np
import pytest

import conftest
import pqt_oracle as po
from pqt_b200 import formats, synth

sys.path.insert(0, os.path.join(conftest.ROOT, "tools", "synthdb"))

pytestmark = pytest.mark.gpu


def test_cuda_generator_matches_numpy():
    import torch
    import synthdb
    mu = synthdb.centres_u8(300, 128, synth.DB_SEED, "cuda")
    assert np.array_equal(mu.cpu().numpy().astype(np.int32), synth.centres(300, 128))
    buf = torch.empty((5000, 128), dtype=torch.uint8, device="cuda")
    for i0 in (0, 123457, 4000000000):
        X = synthdb.db_u8(buf, i0, 5000, mu, synth.DB_SEED).cpu().numpy()
        ids0 = i0 & 0xFFFFFFFF
        ref = synth.db_vectors(ids0, 5000, 128, 300, mu=synth.centres(300, 128)) if i0 < (1 << 32) - 5000 else None
        if ref is not None:
            assert np.array_equal(X, ref)


def test_exact_ground_truth_matches_brute_force():
    import torch
    import synthdb
    n, ncl = 30000, 64
    mu = synthdb.centres_u8(ncl, 128, synth.DB_SEED, "cuda")
    Q8, _ = synthdb.queries_u8(200, n, mu, synth.DB_SEED, synth.QUERY_SEED)
    _, arg = synthdb.exact_1nn(Q8, n, mu, synth.DB_SEED, chunk=4096)
    X = synth.db_vectors(0, n, 128, ncl, mu=synth.centres(ncl, 128)).astype(np.float32)
    gt = po.brute_force_1nn(X, Q8.cpu().numpy().astype(np.float32))
    assert np.array_equal(arg.cpu().numpy().astype(np.uint32), gt)
    # split over two id ranges (what the ranks of a multi-GPU run do)
    s0, a0 = synthdb.exact_1nn(Q8, n, mu, synth.DB_SEED, chunk=4096, i_lo=0, i_hi=n // 2)
    s1, a1 = synthdb.exact_1nn(Q8, n, mu, synth.DB_SEED, chunk=4096, i_lo=n // 2, i_hi=n)
    both = torch.where(s1 < s0, a1, a0)
    assert np.array_equal(both.cpu().numpy().astype(np.uint32), gt)


def test_tool_synthdb_files_match_oracle_builder(tmp_path):
    import synthdb
    synthdb.build()
    n, ncl, dim, p, c1, c2, LP, hs = 20000, 128, 128, 4, 16, 8, 32, 65537
    mu = synth.centres(ncl, dim)
    X = synth.db_vectors(0, n, dim, ncl, mu=mu).astype(np.float32)
    cb1, cb2 = synth.train_tree(X[:4000], p, c1, c2, iters=4)
    base = str(tmp_path / "synth")
    paths = synthdb.index_files(base, dim, p, c1, c2, LP)
    formats.write_ppqt(paths["ppqt"], dim, p, cb1, cb2)
    mu.astype(np.uint8).tofile(str(tmp_path / "mu.u8"))
    synthdb.run_tool(base, n, dim, p, c1, c2, LP, hs, ncl, synth.DB_SEED, str(tmp_path / "mu.u8"),
                     chunksize=7000)
    assert synthdb.files_complete(paths, n, hs, LP)
    prm = po.default_params(dim, p, c1, c2, LP, hash_size=hs)
    ref = po.build_index(prm, cb1, cb2, X, k1_build=16)
    assert np.array_equal(np.fromfile(paths["count"], np.uint32), ref["counts"])
    assert np.array_equal(np.fromfile(paths["prefix"], np.uint32), ref["prefix"])
    assert np.array_equal(np.fromfile(paths["dbIdx"], np.uint32), ref["db_idx"])
    assert np.array_equal(np.fromfile(paths["lines"], np.uint32).reshape(n, LP), ref["lines"])
