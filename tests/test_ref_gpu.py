"""Cross-check against the REFERENCE ITSELF on the GPU box: oracle/_ref/libpqt_ref_gpu.so
is the reference's own pqt/*.cu compiled unmodified for sm_100a (oracle/Makefile,
oracle/ref_gpu_shim.cu).  The oracle (and through tests/test_gpu_parity.py the product)
must reproduce what those kernels compute."""
import os
import subprocess
import sys

import numpy as np
import pytest

import conftest
import pqt_oracle as po
from pqt_b200 import formats, synth

pytestmark = pytest.mark.gpu

REF_LIB = os.path.join(conftest.ROOT, "oracle", "_ref", "libpqt_ref_gpu.so")
HASH = 400000000  # compiled into the reference (pqt/PerturbationProTree.hh:12)
K_BIG = 16        # kVec of the queryBIGKNNRerank2 cross-check


def _run_case(tmp, N, QN, seed):
    dim, p, c1, c2, LP, k1, k = 128, 4, 16, 8, 16, 8, 1024
    mu = synth.centres(256, dim, synth.DB_SEED + seed)
    X = synth.db_vectors(0, N, dim, 256, synth.DB_SEED + seed, mu).astype(np.float32)
    Q = synth.query_vectors(QN, N, dim, 256, synth.DB_SEED + seed, synth.QUERY_SEED + seed,
                            mu)[0].astype(np.float32)
    cb1, cb2 = synth.train_tree(X[:5000], p, c1, c2, iters=6, seed=5)
    ppqt = str(tmp / ("ref%d_128_4_16_8.ppqt" % seed))
    formats.write_ppqt(ppqt, dim, p, cb1, cb2)
    prm = po.default_params(dim, p, c1, c2, LP, hash_size=HASH)
    index = po.build_index(prm, cb1, cb2, X, k1_build=16)
    nz = np.nonzero(index["counts"])[0].astype(np.uint32)
    case = str(tmp / ("case%d.npz" % seed))
    np.savez(case, X=X, Q=Q, dim=dim, p=p, c1=c1, c2=c2, LP=LP, k1=k1, k=k, hash_size=HASH, k_big=K_BIG,
             nz_bins=nz, nz_counts=index["counts"][nz], db_idx=index["db_idx"], lines=index["lines"])
    out = str(tmp / ("out%d.npz" % seed))
    runner = os.path.join(conftest.ROOT, "tests", "ref_gpu_runner.py")
    try:
        r = subprocess.run([sys.executable, runner, case, out, ppqt], capture_output=True,
                           text=True, timeout=600)
    except subprocess.TimeoutExpired:
        r = None
    stages = dict(np.load(out + ".stages.npz")) if os.path.exists(out + ".stages.npz") else None
    full = dict(np.load(out)) if os.path.exists(out) else None
    big = dict(np.load(out + ".big.npz")) if os.path.exists(out + ".big.npz") else None
    if stages is None:
        pytest.fail("the reference kernels did not produce the stage outputs: %s"
                    % (r.stderr[-2000:] if r else "timeout"))
    d0, i0, st0 = po.query_knn(prm, cb1, cb2, index["prefix"], index["counts"], index["db_idx"],
                               index["lines"], Q, k, stages=True)
    return dict(prm=prm, index=index, stages=stages, full=full, big=big, oracle=(d0, i0, st0),
                X=X, Q=Q, cb1=cb1, cb2=cb2, k=k, shape=(dim, p, c1, c2, LP))


@pytest.fixture(scope="module")
def ref_run(tmp_path_factory):
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libpqt_ref_gpu.so was not built (no reference tree at build time)")
    return _run_case(tmp_path_factory.mktemp("refgpu"), 20000, 48, 0)


@pytest.fixture(scope="module")
def ref_run_sparse(tmp_path_factory):
    """few vectors -> short candidate lists: the regime in which the reference's
    rerankKernelFast is free of its shared-memory race (see the end-to-end test)"""
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libpqt_ref_gpu.so was not built (no reference tree at build time)")
    return _run_case(tmp_path_factory.mktemp("refgpu_sparse"), 6000, 96, 9)


def test_reference_kernels_steps_a_to_d(ref_run):
    st, (d0, i0, st0) = ref_run["stages"], ref_run["oracle"]
    assert np.array_equal(st["cb_dist"], ref_run["index"]["cb_dist"])   # computeCBL1L1Dist
    assert np.array_equal(st["assign"], st0["assign"])                  # Step A
    assert np.array_equal(st["lut"], st0["lut"])                        # Step B
    assert np.array_equal(st["assign_val"], st0["assign_val"])          # Step C
    assert np.array_equal(st["assign_idx"], st0["assign_idx"])
    assert np.array_equal(st["n_bins"], st0["n_bins"])                  # Step D
    assert np.array_equal(st["bins"], st0["bins"])


def test_reference_index_build(ref_run):
    st, index = ref_run["stages"], ref_run["index"]
    nz = np.nonzero(index["counts"])[0].astype(np.uint32)
    assert np.array_equal(st["build_nonzero_bins"], nz)                 # buildKBestDB bins
    assert np.array_equal(st["build_counts"], index["counts"][nz])
    assert np.array_equal(st["build_prefix"], index["prefix"][nz])
    # inside a bin the reference's atomicInc order is run dependent: compare as sets
    for b in nz[:500]:
        lo, n = int(index["prefix"][b]), int(index["counts"][b])
        assert sorted(st["build_dbidx"][lo:lo + n]) == list(index["db_idx"][lo:lo + n])
    assert np.array_equal(st["build_lines16"], index["lines"])          # lineDist encoder


# rerankKernelFast (pqt/PerturbationProTree.cu:5189-5351) hands out candidates to groups of
# LP lanes through shared memory without a barrier (laneA[], idx[]: "fetch next a", :5285-5326).
# The first round (blockDim / LP = 64 candidates with the launch shape of rerankKBestVectors)
# is written before any lane reads; from the second round on the other lanes of a group race
# with the writer.  On B200 (independent thread scheduling) exactly the candidates beyond
# the first 64 come out with run-to-run varying distances (observed; SURVEY.md section 5
# predicts it).  So the reference's own output pins: the candidate ids of every query
# (Step E1), and ids + distances + order wherever nVec <= 64.
RACE_FREE = 64


def _check_end_to_end(run):
    if run["full"] is None:
        pytest.xfail("the reference's queryKNN did not complete on this GPU")
    full, (d0, i0, st0), k = run["full"], run["oracle"], run["k"]
    nvec = st0["n_vec"]
    exact = 0
    for q in range(i0.shape[0]):
        n = int(nvec[q])
        # Step E1: same candidates (the reference leaves stale ids in padded slots)
        assert sorted(full["idx"][q, :n]) == sorted(i0[q, :n])
        assert np.all(full["dist"][q, n:] == np.float32(1e7)) and np.all(d0[q, n:] == np.float32(1e7))
        ref_pairs = set(zip(full["dist"][q, :n].tolist(), full["idx"][q, :n].tolist()))
        ora_pairs = set(zip(d0[q, :n].tolist(), i0[q, :n].tolist()))
        if n <= RACE_FREE:
            assert np.array_equal(full["dist"][q], d0[q])
            assert np.array_equal(full["idx"][q, :n], i0[q, :n])
            exact += 1
        else:
            # the race-free first round must still agree
            assert len(ref_pairs & ora_pairs) >= min(len(ora_pairs), RACE_FREE) - (n - len(ora_pairs))
    return exact


def test_reference_query_knn_end_to_end(ref_run):
    _check_end_to_end(ref_run)


def test_reference_query_knn_end_to_end_race_free_regime(ref_run_sparse):
    exact = _check_end_to_end(ref_run_sparse)
    assert exact >= 20  # enough queries compared bit-for-bit against the reference's output


def test_reference_kernels_steps_a_to_d_sparse(ref_run_sparse):
    test_reference_kernels_steps_a_to_d(ref_run_sparse)


def _check_big(run):
    """queryBIGKNNRerank2 on the reference's own kernels: prepare2DDistSequence, the bins of
    getBIGBins2D (bit-exact, in order), and the result ids / race-free distances."""
    if run["big"] is None:
        pytest.xfail("the reference's queryBIGKNNRerank2 did not complete on this GPU")
    big, index = run["big"], run["index"]
    dim, p, c1, c2, LP = run["shape"]
    assert np.array_equal(big["seq2d"], po.dist_seq_2d(512))
    prm = po.big_params(dim, p, c1, c2, LP, hash_size=HASH)
    d0, i0, info = po.query_big_knn_rerank2(prm, run["cb1"], run["cb2"], index["prefix"],
                                            index["counts"], index["db_idx"], index["lines"],
                                            run["Q"], K_BIG)
    # The reference's selectBinKernel2DFinal relies on warp-synchronous scans (scan_warp2,
    # pqt/bitonicSort.cuh:112-132, no __syncwarp): on this GPU its kept-bin count varies from
    # run to run for a few queries (observed: 2 of 48, tools/dbg_ref_gpu.py).  Queries whose bin
    # count agrees are compared strictly; at least 85 % must agree.
    compared = agree = 0
    for q in range(run["Q"].shape[0]):
        if info["ambiguous"][q] or info["ran_off_table"][q]:
            # slope index on a rounding boundary of logf (host/device may differ), or the walk
            # reached the end of d_distSeq: the reference reads past its allocation from there
            continue
        compared += 1
        if big["big_n_bins"][q] != info["n_bins"][q]:
            continue
        n = int(info["n_vec"][q])
        if sorted(big["big_idx"][q, :n]) != sorted(i0[q, :n]):
            continue
        # rerankBIGKernelFast runs max(pow2ceil(k), dim) = 128 threads here: only the first
        # 128 / LP = 8 candidates are evaluated before the shared-memory race sets in
        race_free = max(po.pow2ceil(K_BIG), dim) // LP
        ref_pairs = set(zip(big["big_dist"][q, :n].tolist(), big["big_idx"][q, :n].tolist()))
        ora_pairs = set(zip(d0[q, :n].tolist(), i0[q, :n].tolist()))
        if n <= race_free:
            assert np.array_equal(big["big_dist"][q], d0[q])
            assert np.array_equal(big["big_idx"][q, :n], i0[q, :n])
        else:
            assert len(ref_pairs & ora_pairs) >= min(len(ora_pairs), race_free) - (n - len(ora_pairs))
        agree += 1
    assert compared >= run["Q"].shape[0] // 4
    assert agree >= 0.85 * compared


def test_reference_big_variant(ref_run):
    _check_big(ref_run)


def test_reference_big_variant_sparse(ref_run_sparse):
    _check_big(ref_run_sparse)
