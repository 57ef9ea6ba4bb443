"""Runs the reference's own CUDA kernels (oracle/_ref/libpqt_ref_gpu.so, built from
/root/reference/pqt/*.cu) on a small case and stores what they produce.  Executed in a
subprocess by tests/test_ref_gpu.py so that a hang in the reference's warp-synchronous
code cannot take the test session down.
usage: python tests/ref_gpu_runner.py <case.npz> <out.npz> <ppqt path>"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libpqt_ref_gpu.so")


def main(case_path, out_path, ppqt):
    c = dict(np.load(case_path))
    L = C.CDLL(LIB)
    L.refgpu_create.restype = C.c_void_p
    L.refgpu_hash_size.restype = C.c_uint32
    vp = C.c_void_p
    dim, p, c1, c2, LP, k1, k = (int(c[n]) for n in ("dim", "p", "c1", "c2", "LP", "k1", "k"))
    hs = int(L.refgpu_hash_size())
    assert hs == int(c["hash_size"])
    N, QN = c["X"].shape[0], c["Q"].shape[0]
    h = vp(L.refgpu_create(dim, p))
    L.refgpu_read_tree(h, ppqt.encode())
    out = {}
    # ---- the reference's own index build (buildKBestDB + lineDist, LP = 16)
    prefix = np.zeros(hs, np.uint32)
    counts = np.zeros(hs, np.uint32)
    dbidx = np.zeros(N, np.uint32)
    lines16 = np.zeros((N, 16), np.uint32)
    X = np.ascontiguousarray(c["X"], np.float32)
    L.refgpu_build(h, vp(X.ctypes.data), N, hs, vp(prefix.ctypes.data), vp(counts.ctypes.data),
                   vp(dbidx.ctypes.data), vp(lines16.ctypes.data))
    nz = np.nonzero(counts)[0].astype(np.uint32)
    out.update(build_nonzero_bins=nz, build_counts=counts[nz], build_prefix=prefix[nz],
               build_dbidx=dbidx, build_lines16=lines16)
    L.refgpu_destroy(h)
    # ---- query side on the ORACLE-built index (deterministic dbIdx order)
    h = vp(L.refgpu_create(dim, p))
    L.refgpu_read_tree(h, ppqt.encode())
    prefix[:] = 0
    counts[:] = 0
    counts[c["nz_bins"]] = c["nz_counts"]
    prefix[:] = (np.cumsum(counts, dtype=np.uint64) - counts).astype(np.uint32)
    db_idx = np.ascontiguousarray(c["db_idx"], np.uint32)
    lines = np.ascontiguousarray(c["lines"], np.uint32)
    L.refgpu_set_db(h, N, vp(prefix.ctypes.data), vp(counts.ctypes.data), vp(db_idx.ctypes.data))
    L.refgpu_set_lines(h, vp(lines.ctypes.data), N, LP)
    Q = np.ascontiguousarray(c["Q"], np.float32)
    n = k1 * c2
    max_bins = 4096
    assign = np.zeros((QN, k1, p), np.uint32)
    lut = np.zeros((QN, LP, c1), np.float32)
    aval = np.zeros((QN, p, n), np.float32)
    aidx = np.zeros((QN, p, n), np.uint32)
    bins = np.zeros((QN, max_bins), np.uint32)
    nbins = np.zeros(QN, np.uint32)
    cbd = np.zeros((c1, c1, LP), np.float32)
    L.refgpu_stages(h, vp(Q.ctypes.data), QN, k1, max_bins, vp(assign.ctypes.data),
                    vp(lut.ctypes.data), vp(aval.ctypes.data), vp(aidx.ctypes.data),
                    vp(bins.ctypes.data), vp(nbins.ctypes.data), vp(cbd.ctypes.data))
    out.update(assign=assign, lut=lut, assign_val=aval, assign_idx=aidx, bins=bins, n_bins=nbins,
               cb_dist=cbd)
    np.savez(out_path + ".stages.npz", **out)  # keep the stages even if queryKNN misbehaves
    idx = np.zeros((QN, k), np.uint32)
    dist = np.zeros((QN, k), np.float32)
    L.refgpu_query_knn(h, vp(Q.ctypes.data), QN, k, vp(idx.ctypes.data), vp(dist.ctypes.data))
    out.update(idx=idx, dist=dist)
    np.savez(out_path, **out)
    # ---- the 1-B variant (queryBIGKNNRerank2) with a small k so that the bin walk ends
    # long before the reference would read past its d_distSeq allocation
    kb = int(c["k_big"]) if "k_big" in c else 0
    if kb:
        cap = max(kb, 32)
        bb = np.zeros((QN, cap), np.uint32)
        nb = np.zeros(QN, np.uint32)
        bi = np.zeros((QN, kb), np.uint32)
        bd = np.zeros((QN, kb), np.float32)
        seq2d = np.zeros((10, 65536), np.uint32)
        L.refgpu_big(h, vp(Q.ctypes.data), QN, kb, cap, vp(bb.ctypes.data), vp(nb.ctypes.data),
                     vp(bi.ctypes.data), vp(bd.ctypes.data), vp(seq2d.ctypes.data))
        np.savez(out_path + ".big.npz", big_bins=bb, big_n_bins=nb, big_idx=bi, big_dist=bd,
                 seq2d=seq2d)


if __name__ == "__main__":
    main(*sys.argv[1:4])
