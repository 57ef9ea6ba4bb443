import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "product-quantization-tree_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def make_case(N=20000, QN=64, dim=128, p=4, c1=16, c2=8, LP=16, hash_size=1000003,
              n_clusters=256, seed=0, **over):
    """Small synthetic index built with the ORACLE's builder (test infrastructure)."""
    import pqt_oracle as po
    from pqt_b200 import synth
    mu = synth.centres(n_clusters, dim, synth.DB_SEED + seed)
    X = synth.db_vectors(0, N, dim, n_clusters, synth.DB_SEED + seed, mu).astype(np.float32)
    Qu, src = synth.query_vectors(QN, N, dim, n_clusters, synth.DB_SEED + seed,
                                  synth.QUERY_SEED + seed, mu)
    Q = Qu.astype(np.float32)
    cb1, cb2 = synth.train_tree(X[:min(N, 5000)], p, c1, c2, iters=6, seed=77 + seed)
    prm = po.default_params(dim, p, c1, c2, LP, hash_size=hash_size, **over)
    idx = po.build_index(prm, cb1, cb2, X, k1_build=min(16, c1))
    return dict(prm=prm, X=X, Q=Q, src=src, cb1=cb1, cb2=cb2, **idx)


@pytest.fixture(scope="session")
def case_small():
    return make_case()


@pytest.fixture(scope="session")
def case_lp32():
    return make_case(N=12000, QN=48, LP=32, hash_size=65537, seed=3)
