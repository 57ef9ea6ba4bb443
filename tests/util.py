"""Shared helpers for the test-suite (oracle side + C-ABI side)."""
import os
import zlib

import numpy as np

import conftest  # noqa: F401  (sys.path)
import pqt_oracle as po
from pqt_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def golden_case(name):
    """Rebuilds the inputs of a golden fixture: integer-only synthetic data is
    regenerated, codebooks come from the fixture, the index from the oracle's builder."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    N, QN, dim, p = int(g["N"]), int(g["QN"]), int(g["dim"]), int(g["p"])
    c1, c2, LP, hs, seed = int(g["c1"]), int(g["c2"]), int(g["LP"]), int(g["hash_size"]), int(g["seed"])
    ncl = 256
    mu = synth.centres(ncl, dim, synth.DB_SEED + seed)
    X = synth.db_vectors(0, N, dim, ncl, synth.DB_SEED + seed, mu).astype(np.float32)
    Qu, src = synth.query_vectors(QN, N, dim, ncl, synth.DB_SEED + seed, synth.QUERY_SEED + seed, mu)
    Q = Qu.astype(np.float32)
    prm = po.default_params(dim, p, c1, c2, LP, hash_size=hs)
    idx = po.build_index(prm, g["cb1"], g["cb2"], X, k1_build=min(16, c1))
    case = dict(prm=prm, X=X, Q=Q, src=src, cb1=g["cb1"], cb2=g["cb2"], **idx)
    return case, g


def oracle_query(case, k, stages=False, **over):
    prm = case["prm"]
    if over:
        prm = po.Params.from_buffer_copy(prm)
        for kk, v in over.items():
            setattr(prm, kk, v)
    return po.query_knn(prm, case["cb1"], case["cb2"], case["prefix"], case["counts"],
                        case["db_idx"], case["lines"], case["Q"], k, stages=stages)


def make_gpu_index(case, device=0, shard=None, **params):
    """A PerturbationProTree handle loaded through the C ABI exactly like tool_query does:
    tree -> setDB(host prefix/counts/dbIdx) -> line codes."""
    import pqt_b200
    prm = case["prm"]
    t = pqt_b200.PerturbationProTree(prm.dim, prm.p, prm.p, device)
    kw = dict(k1=prm.k1, max_bins=prm.max_bins, max_trials=prm.max_trials,
              bin_threads=prm.bin_threads, max_vec_per_bin=prm.max_vec_per_bin,
              hash_size=prm.hash_size)
    kw.update(params)
    t.set_params(**kw)
    t.setTree(case["cb1"], case["cb2"].reshape(prm.p, prm.c1, prm.c2, prm.dim // prm.p))
    if shard is not None:
        t.setShard(*shard)
    N = case["db_idx"].size
    t.setDB(N, case["prefix"], case["counts"], case["db_idx"])
    t.setLines(case["lines"], N, prm.line_parts)
    return t
