"""GPU parity: the CUDA path, called through the C ABI, against the oracle on the same
inputs -- bit-exact for ids, bins and every float (the kernels pin the reference's
rounding order, so distances are compared for equality, well inside the 1e-4 relative
tolerance north_star states)."""
import os

import numpy as np
import pytest

import pqt_oracle as po
from util import golden_case, make_gpu_index, oracle_query

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4  # north_star's tolerance for ADC distances; we assert equality below


def _stages(t, prm, QN, k):
    n = prm.k1 * prm.c2
    mv = t.candidateWidth(k)
    return dict(
        assign=t.stage("assign", (QN, prm.k1, prm.p), np.uint32),
        lut=t.stage("lut", (QN, prm.line_parts, prm.c1), np.float32),
        assign_val=t.stage("assign_val", (QN, prm.p, n), np.float32),
        assign_idx=t.stage("assign_idx", (QN, prm.p, n), np.uint32),
        bins=t.stage("bins", (QN, prm.max_bins), np.uint32),
        n_bins=t.stage("n_bins", (QN,), np.uint32),
        select_idx=t.stage("select_idx", (QN, mv), np.uint32),
        n_vec=t.stage("n_vec", (QN,), np.uint32))


def _check_all_stages(case, k, **over):
    d0, i0, st0 = oracle_query(case, k, stages=True, **over)
    t = make_gpu_index(case, **over)
    t.debug(True)
    QN = case["Q"].shape[0]
    i1, d1 = t.queryKNN(case["Q"], QN, k)
    prm = po.Params.from_buffer_copy(case["prm"])
    for kk, v in over.items():
        setattr(prm, kk, v)
    st1 = _stages(t, prm, QN, k)
    seq0, m, _ = po.dist_seq(prm.c2 * prm.k1, prm.p)
    assert np.array_equal(t.stage("dist_seq", (65536,), np.uint32), seq0)
    assert np.array_equal(t.stage("cb_dist", (prm.c1, prm.c1, prm.line_parts), np.float32),
                          case["cb_dist"])
    for name in ("assign", "lut", "assign_val", "n_bins", "bins", "n_vec", "select_idx"):
        assert np.array_equal(st1[name], st0[name]), name
    # assign_idx: padded slots (only when k1*c2 is not a power of two) are unspecified
    assert np.array_equal(st1["assign_idx"], st0["assign_idx"])
    assert np.array_equal(i1, i0)
    assert np.array_equal(d1, d0)
    assert np.allclose(d1, d0, rtol=REL_TOL, atol=0)
    t.close()
    return d1, i1


def test_small_case_every_stage(case_small):
    _check_all_stages(case_small, 256)


def test_lp32_case_every_stage(case_lp32):
    _check_all_stages(case_lp32, 512)


@pytest.mark.parametrize("name", ["c1_16_c2_8_lp16", "c1_32_c2_32_lp32"])
def test_golden_fixtures(name):
    case, g = golden_case(name)
    t = make_gpu_index(case)
    i1, d1 = t.queryKNN(case["Q"], case["Q"].shape[0], int(g["k"]))
    assert np.array_equal(i1, g["idx"])
    assert np.array_equal(d1, g["dist"])
    t.close()


@pytest.mark.parametrize("k", [1, 3, 32, 1000, 4096])
def test_k_sweep(case_small, k):
    d0, i0 = oracle_query(case_small, k)
    t = make_gpu_index(case_small)
    i1, d1 = t.queryKNN(case_small["Q"], case_small["Q"].shape[0], k)
    assert np.array_equal(i1, i0) and np.array_equal(d1, d0)
    t.close()


def test_tied_distances_follow_the_network_order():
    """Duplicate DB vectors give bit-equal ADC distances; the result order of ties is then
    whatever the reference's bitonic network produces (not a stable sort).  Exercises the
    exact-network fallback of the ranking kernels."""
    import conftest
    c = conftest.make_case(N=6000, QN=48, hash_size=65537, seed=11)
    # make every vector appear 4 times: rebuild the index over the duplicated set
    X = np.concatenate([c["X"][:1500]] * 4)
    rng = np.random.default_rng(3)
    X = X[rng.permutation(X.shape[0])]
    idx = po.build_index(c["prm"], c["cb1"], c["cb2"], X, k1_build=16)
    c.update(X=X, **idx)
    for k in (64, 1024, 4096):
        d0, i0 = oracle_query(c, k)
        ties = sum(int(np.sum(np.diff(d0[q][i0[q] != po.PAD_IDX]) == 0)) for q in range(48))
        assert ties > 100  # the case really has ties
        t = make_gpu_index(c)
        i1, d1 = t.queryKNN(c["Q"], 48, k)
        assert np.array_equal(d1, d0)
        assert np.array_equal(i1, i0)
        t.close()


@pytest.fixture(scope="module")
def case_dense():
    """Every query fills its 4096-candidate budget (small hash => no empty bins): the regime
    of the 100M / 1B indexes.  About half of the queries hold bit-equal distances of
    different vectors, a third of the candidates are duplicates of a vector listed twice."""
    import conftest
    return conftest.make_case(N=200000, QN=256, c1=32, c2=32, LP=16, hash_size=20011,
                              n_clusters=1024, seed=5)


@pytest.mark.parametrize("k", [4096, 1000, 37])
def test_dense_candidate_lists_fast_ranking(case_dense, k):
    """rank_mode 0 (composite-key sort + tie handling) and rank_mode 1 (the reference's
    network on every query) both return the oracle's ids and distances."""
    c = case_dense
    QN = c["Q"].shape[0]
    d0, i0 = oracle_query(c, k)
    if k == 4096:
        ties = sum(int(np.any((d0[q][1:] == d0[q][:-1]) & (i0[q][1:] != i0[q][:-1]))) for q in range(QN))
        assert ties > QN // 8  # ties between different vectors are common here
    for mode in (0, 1):
        t = make_gpu_index(c, rank_mode=mode)
        t.profile(True)
        t.reset_stats()
        i1, d1 = t.queryKNN(c["Q"], QN, k)
        st = t.stats()
        assert np.array_equal(d1, d0), "rank_mode %d" % mode
        assert np.array_equal(i1, i0), "rank_mode %d" % mode
        if mode == 0:
            assert st.exact_rank_queries < QN // 8  # the network is the exception, not the rule
            if k == 4096:
                assert st.tie_resolved_queries > QN // 8  # ties re-ordered by the bit-plane simulation
        t.close()


def test_mid_length_candidate_lists_into_poisoned_outputs():
    """Candidate lists of a few hundred to a few thousand entries with k = 4096: every output
    slot, pads included, must be written (the outputs start out poisoned).  Covers the list
    lengths at which every thread of a group sorts (no idle thread left for the pads)."""
    import conftest
    import torch
    c = conftest.make_case(N=120000, QN=192, c1=32, c2=32, LP=16, hash_size=3000017,
                           n_clusters=512, seed=9)
    QN, k = c["Q"].shape[0], 4096
    d0, i0 = oracle_query(c, k)
    nv = (i0 != po.PAD_IDX).sum(1)
    assert ((nv > 256) & (nv <= 1024)).any() and ((nv > 1024) & (nv <= 2048)).any(), np.percentile(nv, [0, 25, 50, 75, 100])
    t = make_gpu_index(c)
    Qd = torch.from_numpy(c["Q"]).cuda()
    for poison in (0x7FC00000, 0):
        oi = torch.full((QN, k), 0x12345678, dtype=torch.int32, device="cuda")
        od = torch.full((QN, k), poison, dtype=torch.int32, device="cuda").view(torch.float32)
        t.queryKNN(Qd, QN, k, oi, od)
        assert np.array_equal(od.cpu().numpy(), d0)
        assert np.array_equal(oi.cpu().numpy().view(np.uint32), i0)
    t.close()


def test_distances_at_or_above_the_pad_value(case_small):
    """Distances >= 1e7 sort behind / among the 1e7 padding in the reference's network."""
    c = dict(case_small)
    c["Q"] = (case_small["Q"] * 40.0).astype(np.float32)  # far away queries: huge distances
    d0, i0 = oracle_query(c, 512)
    assert (d0[i0 != po.PAD_IDX] >= 1e7).any()
    t = make_gpu_index(c)
    i1, d1 = t.queryKNN(c["Q"], c["Q"].shape[0], 512)
    assert np.array_equal(d1, d0) and np.array_equal(i1, i0)
    t.close()


def test_truncation_budgets(case_small):
    _check_all_stages(case_small, 64, max_bins=8, max_vec_per_bin=3)
    _check_all_stages(case_small, 64, max_trials=1)
    _check_all_stages(case_small, 128, bin_threads=256, max_trials=7, k1=4)
    # probe budgets that end inside / span several 16384-code visiting blocks (bins4_kernel)
    _check_all_stages(case_small, 256, max_trials=5)
    _check_all_stages(case_small, 256, max_trials=20)
    _check_all_stages(case_small, 256, max_trials=64)


def test_empty_index_and_single_query(case_small):
    c = dict(case_small)
    c["counts"] = np.zeros_like(case_small["counts"])
    c["prefix"] = np.zeros_like(case_small["prefix"])
    t = make_gpu_index(c)
    i1, d1 = t.queryKNN(c["Q"], c["Q"].shape[0], 32)
    assert np.all(i1 == po.PAD_IDX) and np.all(d1 == np.float32(1e7))
    t.close()
    c = dict(case_small)
    c["Q"] = case_small["Q"][5:6]
    d0, i0 = oracle_query(c, 256)
    t = make_gpu_index(c)
    i1, d1 = t.queryKNN(c["Q"], 1, 256)
    assert np.array_equal(i1, i0) and np.array_equal(d1, d0)
    t.close()


def test_device_pointers_and_max_vec_extension(case_small):
    import torch
    c = case_small
    QN = c["Q"].shape[0]
    d0, i0 = oracle_query(c, 1024)
    t = make_gpu_index(c, max_vec=1024)
    Qd = torch.from_numpy(c["Q"]).cuda()
    oi = torch.zeros((QN, 10), dtype=torch.int32, device="cuda")
    od = torch.zeros((QN, 10), dtype=torch.float32, device="cuda")
    t.queryKNN(Qd, QN, 10, oi, od)
    # first k of the width-1024 ranking (params.max_vec extension)
    assert np.array_equal(oi.cpu().numpy().view(np.uint32), i0[:, :10])
    assert np.array_equal(od.cpu().numpy(), d0[:, :10])
    # pinned host outputs
    pi = torch.zeros((QN, 10), dtype=torch.int32).pin_memory()
    pd = torch.zeros((QN, 10), dtype=torch.float32).pin_memory()
    t.queryKNN(Qd, QN, 10, pi, pd)
    assert np.array_equal(pi.numpy().view(np.uint32), i0[:, :10])
    t.close()


def test_error_behaviour(case_small, tmp_path):
    import pqt_b200
    c = case_small
    prm = c["prm"]
    t = pqt_b200.PerturbationProTree(prm.dim, prm.p)
    with pytest.raises(pqt_b200.PqtError):  # query before anything is loaded
        t.queryKNN(c["Q"], 4, 4)
    t.set_params(hash_size=prm.hash_size)
    t.setTree(c["cb1"], c["cb2"].reshape(prm.p, prm.c1, prm.c2, -1))
    t.setDB(c["db_idx"].size, c["prefix"], c["counts"], c["db_idx"])
    with pytest.raises(pqt_b200.PqtError) as e:  # the shipped tool_query forgets the lines
        t.queryKNN(c["Q"], 4, 4)
    assert "line codes" in str(e.value)
    with pytest.raises(pqt_b200.PqtError):  # lineparts must be a power of two <= 32
        t.setLines(np.zeros((c["db_idx"].size, 24), np.uint32), c["db_idx"].size, 24)
    t.setLines(c["lines"], c["db_idx"].size, prm.line_parts)
    with pytest.raises(pqt_b200.PqtError):
        t.queryKNN(c["Q"], 4, 8192)  # candidate width > 4096
    with pytest.raises(pqt_b200.PqtError):
        t.readTreeFromFile(str(tmp_path / "missing.ppqt"))
    # tree file round trip through the reference's .ppqt format
    path = str(tmp_path / "t_128_4_16_8.ppqt")
    t.writeTreeToFile(path)
    from pqt_b200 import formats
    f = formats.read_ppqt(path)
    assert np.array_equal(f["cb1"], c["cb1"]) and f["c2"] == prm.c2
    t2 = pqt_b200.PerturbationProTree(8, 1)  # constructor args are overridden by the file
    t2.readTreeFromFile(path)
    assert t2.shape() == (prm.dim, prm.p, prm.c1, prm.c2)
    t.close()
    t2.close()


def test_gpu_builder_matches_oracle_builder(case_small):
    import pqt_b200
    c = case_small
    prm = c["prm"]
    t = pqt_b200.PerturbationProTree(prm.dim, prm.p)
    t.set_params(hash_size=prm.hash_size, k1_build=min(16, prm.c1))
    t.setTree(c["cb1"], c["cb2"].reshape(prm.p, prm.c1, prm.c2, -1))
    N = c["X"].shape[0]
    t.buildKBestDB(c["X"], N)
    t.lineDist(c["X"], N, prm.line_parts)
    prefix, counts, db_idx = t.getDB()
    assert np.array_equal(counts, c["counts"])
    assert np.array_equal(prefix, c["prefix"])
    assert np.array_equal(db_idx, c["db_idx"])
    assert np.array_equal(t.getLine(), c["lines"])
    d0, i0 = oracle_query(c, 256)
    i1, d1 = t.queryKNN(c["Q"], c["Q"].shape[0], 256)
    assert np.array_equal(i1, i0) and np.array_equal(d1, d0)
    t.close()


def test_loaded_db_round_trips(case_small):
    t = make_gpu_index(case_small)
    prefix, counts, db_idx = t.getDB()
    assert np.array_equal(counts, case_small["counts"])
    assert np.array_equal(prefix, case_small["prefix"])
    assert np.array_equal(db_idx, case_small["db_idx"])
    assert np.array_equal(t.getLine(), case_small["lines"])
    t.close()


def _run_sharded(c, k, world, QN, **params):
    """Bin-range shards in one process: `world` handles on one device wired with plain device
    pointers instead of CUDA-IPC mappings; the three phases run rank after rank (the barriers
    of the multi-process path).  Returns (idx, dist) of all queries."""
    import torch
    Qd = torch.from_numpy(np.ascontiguousarray(c["Q"][:QN])).cuda()
    ts = [make_gpu_index(c, shard=(r, world), **params) for r in range(world)]
    mv = ts[0].candidateWidth(k)
    per = QN // world
    for t in ts:
        t.shardExchangeAlloc(per, mv)
    ptrs = [t.shardExchangePtrs() for t in ts]
    for t in ts:
        t.shardExchangeSetPeers([p[0] for p in ptrs], [p[1] for p in ptrs], [p[2] for p in ptrs])
    for r, t in enumerate(ts):
        t.shardDispatch(Qd, QN, k, r * per, (r + 1) * per)
    torch.cuda.synchronize()  # the cross-rank barrier: every inbox is complete
    for t in ts:
        t.shardScanP2P(QN, k)
    torch.cuda.synchronize()  # the cross-rank barrier: every distance has arrived
    oi = torch.zeros((QN, k), dtype=torch.int32, device="cuda")
    od = torch.zeros((QN, k), dtype=torch.float32, device="cuda")
    for r, t in enumerate(ts):
        t.shardRank(per, k, oi[r * per:(r + 1) * per], od[r * per:(r + 1) * per])
    for t in ts:
        t.close()
    return oi.cpu().numpy().view(np.uint32), od.cpu().numpy()


def test_compact_index_file_round_trip(case_small, case_lp32, tmp_path):
    """pqt_save_index / pqt_load_index: a fresh handle with the same tree answers from the file
    exactly like the handle that wrote it and still expands the reference's arrays; sharded
    handles write and read their own slice; mismatching trees / shards / files are refused"""
    import pqt_b200
    for name, c in (("small", case_small), ("lp32", case_lp32)):
        prm = c["prm"]
        QN = c["Q"].shape[0]
        d0, i0 = oracle_query(c, 256)
        path = str(tmp_path / (name + ".pqtx"))
        t = make_gpu_index(c)
        t.saveIndex(path)
        t.close()
        t2 = pqt_b200.PerturbationProTree(prm.dim, prm.p)
        t2.set_params(hash_size=prm.hash_size, k1=prm.k1, max_bins=prm.max_bins, max_trials=prm.max_trials,
                      bin_threads=prm.bin_threads, max_vec_per_bin=prm.max_vec_per_bin)
        t2.setTree(c["cb1"], c["cb2"].reshape(prm.p, prm.c1, prm.c2, prm.dim // prm.p))
        t2.loadIndex(path)
        i1, d1 = t2.queryKNN(c["Q"], QN, 256)
        assert np.array_equal(i1, i0) and np.array_equal(d1, d0)
        prefix, counts, db_idx = t2.getDB()
        assert np.array_equal(prefix, c["prefix"]) and np.array_equal(counts, c["counts"])
        assert np.array_equal(db_idx, c["db_idx"]) and np.array_equal(t2.getLine(), c["lines"])
        t2.close()
    # a shard's file holds its slice of the codes and refuses another shard's handle
    c = case_small
    prm = c["prm"]
    N = c["db_idx"].size
    path = str(tmp_path / "shard1.pqtx")
    t = make_gpu_index(c, shard=(1, 3))
    t.saveIndex(path)
    t.close()
    lo, hi = N // 3, (2 * N) // 3
    assert os.path.getsize(path) < os.path.getsize(str(tmp_path / "small.pqtx"))
    for shard, ok in (((1, 3), True), ((0, 3), False), (None, False)):
        t3 = pqt_b200.PerturbationProTree(prm.dim, prm.p)
        t3.set_params(hash_size=prm.hash_size)
        t3.setTree(c["cb1"], c["cb2"].reshape(prm.p, prm.c1, prm.c2, prm.dim // prm.p))
        if shard:
            t3.setShard(*shard)
        if ok:
            t3.loadIndex(path)
            assert np.array_equal(t3.getCodesBinOrder(0, hi - lo), c["lines"][c["db_idx"]][lo:hi])
        else:
            with pytest.raises(pqt_b200.PqtError):
                t3.loadIndex(path)
        t3.close()
    # wrong tree shape, wrong hash size, truncated file, not an index file
    t4 = pqt_b200.PerturbationProTree(prm.dim, prm.p)
    t4.set_params(hash_size=prm.hash_size + 2)
    t4.setTree(c["cb1"], c["cb2"].reshape(prm.p, prm.c1, prm.c2, prm.dim // prm.p))
    with pytest.raises(pqt_b200.PqtError):
        t4.loadIndex(str(tmp_path / "small.pqtx"))
    t4.set_params(hash_size=prm.hash_size)
    blob = open(str(tmp_path / "small.pqtx"), "rb").read()
    open(str(tmp_path / "cut.pqtx"), "wb").write(blob[:len(blob) // 2])
    open(str(tmp_path / "junk.pqtx"), "wb").write(b"x" * 4096)
    for bad in ("cut.pqtx", "junk.pqtx", "missing.pqtx"):
        with pytest.raises(pqt_b200.PqtError):
            t4.loadIndex(str(tmp_path / bad))
    with pytest.raises(pqt_b200.PqtError):
        t4.queryKNN(c["Q"], 4, 16)  # a refused file leaves no half-loaded index behind
    t4.loadIndex(str(tmp_path / "small.pqtx"))
    i1, d1 = t4.queryKNN(c["Q"], c["Q"].shape[0], 256)
    d0, i0 = oracle_query(c, 256)
    assert np.array_equal(i1, i0) and np.array_equal(d1, d0)
    t4.close()


@pytest.mark.parametrize("which,world", [("small", 3), ("lp32", 2), ("dense", 3), ("dense", 1)])
def test_shards_in_one_process(which, world, case_small, case_lp32, case_dense):
    """The multi-GPU pipeline (dispatch -> inbox scan with stores into the owner's arrays ->
    ranking) returns the single-index result, ties included; world = 1 is the degenerate case
    where every candidate stays at home."""
    c = {"small": case_small, "lp32": case_lp32, "dense": case_dense}[which]
    k = 4096 if which == "dense" else 256
    QN = (c["Q"].shape[0] // world) * world
    c = dict(c)
    c["Q"] = c["Q"][:QN]
    d0, i0 = oracle_query(c, k)
    for mode in ((0, 1) if which == "dense" else (0,)):
        i1, d1 = _run_sharded(c, k, world, QN, rank_mode=mode)
        assert np.array_equal(d1, d0), "rank_mode %d" % mode
        assert np.array_equal(i1, i0), "rank_mode %d" % mode


def test_sharded_handle_rejects_the_single_gpu_call(case_small):
    import pqt_b200
    t = make_gpu_index(case_small, shard=(0, 2))
    with pytest.raises(pqt_b200.PqtError):
        t.queryKNN(case_small["Q"], 4, 16)
    with pytest.raises(pqt_b200.PqtError):
        t.shardScanP2P(4, 16)  # exchange buffers not connected
    t.close()


def test_eight_parts_take_the_generic_bin_walk():
    """p = 8 (the GIST-style tree shape) with a traversal width of 8 (k1 * c2 = 8, 8^8 codes):
    bins2_kernel instead of the p <= 4 kernels, generic table kernel, LP = 16"""
    import conftest
    c = conftest.make_case(N=6000, QN=16, p=8, c1=8, c2=4, LP=16, hash_size=65537, seed=13, k1=2)
    QN = c["Q"].shape[0]
    t = make_gpu_index(c)
    for k in (16, 256):
        d0, i0 = oracle_query(c, k)
        i1, d1 = t.queryKNN(c["Q"], QN, k)
        assert np.array_equal(d1, d0)
        assert np.array_equal(i1, i0)
    t.close()


def test_traversal_tables_beyond_the_supported_size_are_rejected():
    """p = 8 with the default traversal width 16 would need 16^8 codes (the reference's uint
    nVec wraps to 0 there, pqt/ProTree.cu:139): refused with an error, not answered wrongly"""
    import conftest
    import pqt_b200
    c = conftest.make_case(N=3000, QN=4, p=8, c1=8, c2=4, LP=16, hash_size=65537, seed=13, k1=2)
    t = make_gpu_index(c, k1=8)  # k1 * c2 = 32 -> traversal width 16
    with pytest.raises(pqt_b200.PqtError):
        t.queryKNN(c["Q"], 4, 16)
    t.close()


# ---- a11: the 1-B variant queryBIGKNNRerank2 ---------------------------------------------

def _big_check(c, hash_size, ks):
    prm = c["prm"]
    QN = c["Q"].shape[0]
    bp = po.big_params(prm.dim, prm.p, prm.c1, prm.c2, prm.line_parts, hash_size=hash_size)
    t = make_gpu_index(c)
    t.debug(True)
    for k in ks:
        d0, i0, info = po.query_big_knn_rerank2(bp, c["cb1"], c["cb2"], c["prefix"], c["counts"],
                                                c["db_idx"], c["lines"], c["Q"], k)
        i1, d1 = t.queryBIGKNNRerank2(c["Q"], QN, k)
        assert np.array_equal(t.stage("dist_seq_2d", (10, 65536), np.uint32), po.dist_seq_2d(512))
        ok = ~info["ambiguous"]  # slope index on a logf rounding boundary: host/device may differ
        assert ok.sum() >= QN - 2
        assert np.array_equal(t.stage("n_vec", (QN,), np.uint32)[ok], info["n_vec"][ok])
        assert np.array_equal(t.stage("big_n_bins", (QN,), np.uint32)[ok], info["n_bins"][ok])
        assert np.array_equal(d1[ok], d0[ok])
        assert np.array_equal(i1[ok], i0[ok])
    t.close()


def test_big_variant_dense_index():
    import conftest
    c = conftest.make_case(N=30000, QN=24, c1=16, c2=8, LP=16, hash_size=20011, seed=21)
    _big_check(c, 20011, (16, 256, 1024, 4096))


def test_big_variant_sparse_index(case_small):
    # almost all hash bins empty: hundreds of merge rounds per query, up to the point where
    # the reference would leave its d_distSeq allocation
    c = dict(case_small)
    c["Q"] = case_small["Q"][:12]
    _big_check(c, case_small["prm"].hash_size, (64, 512))


def test_big_variant_rejects_unsupported_shapes(case_small):
    import pqt_b200
    t = make_gpu_index(case_small, big_k1=4)  # 4 * c2 = 32 < 64 sorted entries per part
    with pytest.raises(pqt_b200.PqtError):
        t.queryBIGKNNRerank2(case_small["Q"], 4, 16)
    t.close()
