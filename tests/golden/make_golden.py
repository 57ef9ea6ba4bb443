"""Records the golden end-to-end fixture from the oracle.

The reference ships no golden vectors, so end-to-end parity is
pinned by this recording: codebooks (k-means output is BLAS dependent, so they are
stored, not regenerated), the synthetic-data recipe (integer only, regenerated), and
the oracle's outputs.  Run:  python tests/golden/make_golden.py
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import conftest  # noqa: E402  (sets sys.path for oracle + package)
import pqt_oracle as po  # noqa: E402

CASES = {
    # cpu_version-shaped config C1 (p=4, c1=16, c2=8), small N
    "c1_16_c2_8_lp16": dict(N=20000, QN=64, c1=16, c2=8, LP=16, hash_size=1000003, k=256),
    # bench-shaped codebooks (c1=c2=32), LP=32, tiny hash -> heavy collisions, max_bins hit
    "c1_32_c2_32_lp32": dict(N=8000, QN=32, c1=32, c2=32, LP=32, hash_size=4099, k=1024, seed=5),
}


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def record(name, N, QN, c1, c2, LP, hash_size, k, seed=0):
    c = conftest.make_case(N=N, QN=QN, c1=c1, c2=c2, LP=LP, hash_size=hash_size, seed=seed)
    d, i, st = po.query_knn(c["prm"], c["cb1"], c["cb2"], c["prefix"], c["counts"], c["db_idx"],
                            c["lines"], c["Q"], k, stages=True)
    out = dict(N=N, QN=QN, dim=128, p=4, c1=c1, c2=c2, LP=LP, hash_size=hash_size, k=k, seed=seed,
               cb1=c["cb1"], cb2=c["cb2"], dist=d, idx=i, n_bins=st["n_bins"], n_vec=st["n_vec"],
               assign=st["assign"], crc_lut=crc(st["lut"]), crc_assign_idx=crc(st["assign_idx"]),
               crc_bins=crc(st["bins"]), crc_select_idx=crc(st["select_idx"]),
               crc_lines=crc(c["lines"]), crc_db_idx=crc(c["db_idx"]), crc_counts=crc(c["counts"]),
               crc_X=crc(c["X"]), crc_Q=crc(c["Q"]))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "nvec mean", st["n_vec"].mean(), "nbins mean", st["n_bins"].mean())


if __name__ == "__main__":
    for name, kw in CASES.items():
        record(name, **kw)
