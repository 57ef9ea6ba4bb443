"""The oracle against its recorded golden fixtures and against size-independent
properties of the queryKNN chain."""
import ctypes as C

import numpy as np
import pytest

import pqt_oracle as po
from util import crc, golden_case, oracle_query


@pytest.mark.parametrize("name", ["c1_16_c2_8_lp16", "c1_32_c2_32_lp32"])
def test_golden_end_to_end(name):
    case, g = golden_case(name)
    assert crc(case["X"]) == int(g["crc_X"]) and crc(case["Q"]) == int(g["crc_Q"])
    assert crc(case["lines"]) == int(g["crc_lines"])
    assert crc(case["db_idx"]) == int(g["crc_db_idx"])
    assert crc(case["counts"]) == int(g["crc_counts"])
    d, i, st = oracle_query(case, int(g["k"]), stages=True)
    assert np.array_equal(i, g["idx"])
    assert np.array_equal(d, g["dist"])
    assert np.array_equal(st["n_bins"], g["n_bins"]) and np.array_equal(st["n_vec"], g["n_vec"])
    assert np.array_equal(st["assign"], g["assign"])
    assert crc(st["lut"]) == int(g["crc_lut"])
    assert crc(st["assign_idx"]) == int(g["crc_assign_idx"])
    assert crc(st["bins"]) == int(g["crc_bins"])
    assert crc(st["select_idx"]) == int(g["crc_select_idx"])


def test_dist_seq_properties():
    # prepareDistSequence (pqt/ProTree.cu:128-207): first entry is the origin, codes are a
    # permutation of [0, m^p), sorted by (sum of sqrt(rank), code)
    seq, m, nv = po.dist_seq(8 * 8, 4)
    assert m == 16 and nv == 65536
    assert seq[0] == 0
    assert np.array_equal(np.sort(seq), np.arange(65536, dtype=np.uint32))
    r = np.stack([(seq // 16 ** j) % 16 for j in range(4)], 1).astype(np.float32)
    key = np.zeros(len(seq), np.float32)
    for j in range(4):
        key = (key + np.sqrt(r[:, j])).astype(np.float32)
    assert np.all(np.diff(key) >= 0)
    ties = np.diff(key) == 0
    assert np.all(np.diff(seq.astype(np.int64))[ties] > 0)
    # the 4 single-step neighbours come right after the origin, lowest code first
    assert list(seq[1:5]) == [1, 16, 256, 4096]
    # small trees: m = c2*k1 < 16, unused tail zero
    seq, m, nv = po.dist_seq(3, 2)
    assert m == 3 and nv == 9 and np.all(seq[9:] == 0)
    assert sorted(seq[:9]) == list(range(9))


def test_seg_dist_is_the_pairwise_tree():
    rng = np.random.default_rng(1)
    for n in (2, 8, 32, 128):
        q = rng.uniform(0, 255, n).astype(np.float32)
        c = rng.uniform(0, 255, n).astype(np.float32)
        s = ((q - c).astype(np.float32) ** 2).astype(np.float32)
        stride = n // 2
        while stride:
            s[:stride] = (s[:stride] + s[stride:2 * stride]).astype(np.float32)
            stride //= 2
        assert po.seg_dist(q, c) == s[0]


def test_chain_properties(case_small):
    c = case_small
    prm = c["prm"]
    k = 512
    d, i, st = oracle_query(c, k, stages=True)
    QN = c["Q"].shape[0]
    # ascending distances (test/test1B.cpp:1256-1265)
    assert np.all(np.diff(d, axis=1) >= 0)
    # bins: slot 0 stays 0, nBins <= max_bins, every listed bin (beyond slot 0) is non-empty
    assert np.all(st["bins"][:, 0] == 0)
    assert np.all(st["n_bins"] <= prm.max_bins)
    for q in range(QN):
        nb = st["n_bins"][q]
        listed = st["bins"][q, 1:nb]
        assert np.all(c["counts"][listed] > 0)
        # Step E1 == concatenation of the listed bins' vectors, truncated
        want = []
        for b in st["bins"][q, :nb]:
            n = min(int(c["counts"][b]), prm.max_vec_per_bin)
            want.extend(c["db_idx"][c["prefix"][b]:c["prefix"][b] + n])
        want = np.array(want[:k], np.uint32)
        nv = st["n_vec"][q]
        assert nv == len(want)
        assert np.array_equal(st["select_idx"][q, :nv], want)
        assert np.all(st["select_idx"][q, nv:] == 0)
        # results: a permutation of the candidates + padding
        assert sorted(i[q, :nv]) == sorted(want)
        assert np.all(i[q, nv:] == po.PAD_IDX) and np.all(d[q, nv:] == np.float32(1e7))
    # Step A: assign[q][0][part] is the nearest L1 centroid by the tree-ordered distance
    vl = prm.dim // prm.p
    for q in range(0, QN, 7):
        for part in range(prm.p):
            dd = [po.seg_dist(c["Q"][q, part * vl:(part + 1) * vl],
                              c["cb1"][cc, part * vl:(part + 1) * vl]) for cc in range(prm.c1)]
            assert dd[st["assign"][q, 0, part]] == min(dd)
    # Step C lists are sorted and idx = l2 + l1*c2 with l1 among the k1 cells
    assert np.all(np.diff(st["assign_val"], axis=2) >= 0)
    l1 = st["assign_idx"] // prm.c2
    for q in range(0, QN, 5):
        for part in range(prm.p):
            assert set(l1[q, part]) == set(st["assign"][q, :, part])
    # ADC distance of the winner equals a direct evaluation
    cbd = c["cb_dist"]
    for q in range(0, QN, 9):
        v = po.line_adc(prm, st["lut"][q], cbd, c["lines"][i[q, 0]])
        assert v == d[q, 0]


def test_query_order_and_batching_do_not_matter(case_small):
    c = dict(case_small)
    d0, i0 = oracle_query(c, 128)
    perm = np.random.default_rng(0).permutation(c["Q"].shape[0])
    c["Q"] = case_small["Q"][perm]
    d1, i1 = oracle_query(c, 128)
    assert np.array_equal(i1, i0[perm]) and np.array_equal(d1, d0[perm])
    c["Q"] = case_small["Q"][:1]
    d2, i2 = oracle_query(c, 128)
    assert np.array_equal(i2[0], i0[0])


def test_truncation_rules(case_small):
    # tiny budgets exercise max_bins / max_vec_per_bin / max_vec truncation and the
    # "last kept bin is dropped" off-by-one
    c = case_small
    d, i, st = oracle_query(c, 64, stages=True, max_bins=8, max_vec_per_bin=3)
    assert np.all(st["n_bins"] <= 8)
    for q in range(c["Q"].shape[0]):
        nb = st["n_bins"][q]
        tot = sum(min(3, int(c["counts"][b])) for b in st["bins"][q, :nb])
        assert st["n_vec"][q] == min(64, tot)
    # one trial only: at most bin_threads probes
    d, i, st = oracle_query(c, 64, stages=True, max_trials=1)
    assert np.all(st["n_bins"] <= c["prm"].bin_threads)


def test_empty_and_ragged_inputs(case_small):
    c = dict(case_small)
    # an index whose hash table is empty: every result slot is padding
    c["counts"] = np.zeros_like(case_small["counts"])
    c["prefix"] = np.zeros_like(case_small["prefix"])
    d, i, st = oracle_query(c, 32, stages=True)
    assert np.all(st["n_vec"] == 0) and np.all(st["n_bins"] == 0)
    assert np.all(i == po.PAD_IDX) and np.all(d == np.float32(1e7))
    # k = 1 and a non-power-of-two k (candidate width = pow2ceil(k))
    d1, i1 = oracle_query(case_small, 1)
    d3, i3 = oracle_query(case_small, 3)
    d4, i4 = oracle_query(case_small, 4)
    assert np.array_equal(i3, i4[:, :3]) and np.array_equal(d3, d4[:, :3])
    assert d1.shape == (case_small["Q"].shape[0], 1)
    # unsupported shapes are rejected, not mis-computed
    bad = po.default_params(120, 4, 16, 8, 16)  # dim/p = 30: not a power of two
    with pytest.raises(ValueError):
        po.query_knn(bad, np.zeros((16, 120), np.float32), np.zeros((4, 16, 8, 30), np.float32),
                     c["prefix"], c["counts"], c["db_idx"], c["lines"],
                     np.zeros((1, 120), np.float32), 4)


def test_builder_properties(case_small):
    c = case_small
    prm = c["prm"]
    N = c["X"].shape[0]
    # inverted lists: exclusive prefix of counts, ids ascending inside every bin
    assert c["counts"].sum() == N
    assert np.array_equal(c["prefix"], np.cumsum(c["counts"], dtype=np.uint64).astype(np.uint32)
                          - c["counts"])
    assert np.array_equal(np.sort(c["db_idx"]), np.arange(N, dtype=np.uint32))
    ne = np.nonzero(c["counts"])[0]
    for b in ne[:200]:
        seg = c["db_idx"][c["prefix"][b]:c["prefix"][b] + c["counts"][b]]
        assert np.all(np.diff(seg.astype(np.int64)) > 0)
        assert np.all(c["bin_of"][seg] == b)
    # line codes: p1 != p2, both valid centroids; reconstruction error of the code is the
    # best over all ordered pairs (spot check)
    p1 = c["lines"] & 0xFF
    p2 = (c["lines"] >> 8) & 0xFF
    assert np.all(p1 < prm.c1) and np.all(p2 < prm.c1) and np.all(p1 != p2)
    # a DB vector used as its own query is found in its own bin with a small ADC distance
    c2 = dict(c)
    c2["Q"] = c["X"][:32]
    d, i = oracle_query(c2, 64)
    hit = [q in i[q] for q in range(32)]
    assert np.mean(hit) > 0.9


# ---- a11: the 1-B variant (queryBIGKNNRerank2) -------------------------------------------

def test_dist_seq_2d_properties():
    seq = po.dist_seq_2d(512)
    assert seq.shape == (10, 65536)
    for s in range(10):
        assert seq[s, 0] == 0                      # origin first
        assert len(np.unique(seq[s])) == 65536     # distinct codes of the 512 x 512 grid
        x, y = seq[s] % 512, seq[s] // 512
        sl = np.float32(pow(0.9 * float(np.float32(1.2)), s - 5))
        key = (np.power(x.astype(np.float32), np.float32(0.8)) +
               sl * np.power(y.astype(np.float32), np.float32(0.8)))
        assert np.all(np.diff(key.astype(np.float64)) >= -1e-3)   # sorted by the anisotropic cost
    # flatter slopes walk further along x before stepping in y
    first_y = [int(np.argmax(seq[s] // 512 > 0)) for s in range(10)]
    assert first_y == sorted(first_y)


def test_slope_idx_cases():
    v = np.arange(64, dtype=np.float32)
    assert po.slope_idx(v, v, 256)[0] == 5                        # equal growth -> middle slope
    assert po.slope_idx(v, 1.2 ** 3 * v, 256)[0] == 8
    assert po.slope_idx(v, 100 * v, 256)[0] == 9                  # clamped
    assert po.slope_idx(100 * v, v, 256)[0] == 0
    assert po.slope_idx(np.zeros(64, np.float32), v, 256)[0] == 9   # division by zero -> +inf
    assert po.slope_idx(np.zeros(64, np.float32), np.zeros(64, np.float32), 256)[0] == 0  # NaN


def _dense_case():
    """c1 = c2 = 8, tiny hash: most hash bins are occupied, the regime of the 1-B path"""
    import conftest
    return conftest.make_case(N=30000, QN=24, c1=16, c2=8, LP=16, hash_size=20011, seed=21)


def test_big_path_properties():
    c = _dense_case()
    prm = po.big_params(128, 4, 16, 8, 16, hash_size=20011)
    k = 256
    d, i, info = po.query_big_knn_rerank2(prm, c["cb1"], c["cb2"], c["prefix"], c["counts"],
                                          c["db_idx"], c["lines"], c["Q"], k)
    assert np.all(np.diff(d, axis=1) >= 0)          # test/test1B.cpp:1256-1265
    assert np.all(info["n_vec"] <= k) and np.all(info["n_bins"] <= prm.max_bins)
    assert info["n_vec"].min() > 0
    cbd = c["cb_dist"]
    d0, i0, st = oracle_query(c, k, stages=True)    # same LUT (Step B does not depend on k1)
    for q in range(c["Q"].shape[0]):
        nv = int(info["n_vec"][q])
        assert np.all(i[q, nv:] == po.PAD_IDX)
        real = i[q, :nv]
        assert np.all(real < c["X"].shape[0])
        # reported distances are the ADC distances of the reported ids
        for j in (0, nv // 2, nv - 1):
            assert d[q, j] == po.line_adc(c["prm"], st["lut"][q], cbd, c["lines"][real[j]])
    # a DB vector used as query finds itself
    c2 = dict(c)
    c2["Q"] = c["X"][:16]
    d, i, info = po.query_big_knn_rerank2(prm, c["cb1"], c["cb2"], c["prefix"], c["counts"],
                                          c["db_idx"], c["lines"], c2["Q"], k)
    assert np.mean([q in i[q] for q in range(16)]) > 0.8
