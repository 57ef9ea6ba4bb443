"""ctypes view of the CPU oracle (oracle/libpqt_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpqt_oracle.so")
REF_HOST_PATH = os.path.join(HERE, "_ref", "libpqt_ref_host.so")

NUM_DISTSEQ = 65536
PAD_IDX = 0xFFFFFFFF


class Params(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "dim", "p", "c1", "c2", "line_parts", "k1", "max_bins", "max_trials",
        "bin_threads", "max_vec_per_bin", "hash_size")]


class Stages(C.Structure):
    _fields_ = [("assign", C.c_void_p), ("lut", C.c_void_p), ("assign_val", C.c_void_p),
                ("assign_idx", C.c_void_p), ("bins", C.c_void_p), ("n_bins", C.c_void_p),
                ("select_idx", C.c_void_p), ("n_vec", C.c_void_p)]


def build(force=False):
    """make -C oracle (also builds oracle/_ref when /root/reference exists)."""
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "pqt_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "--no-print-directory"],
                              stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.pqto_pow2ceil.restype = C.c_uint32
        L.pqto_pow2ceil.argtypes = [C.c_uint32]
        L.pqto_to_ushort.restype = C.c_uint16
        L.pqto_to_ushort.argtypes = [C.c_float]
        L.pqto_to_float.restype = C.c_float
        L.pqto_to_float.argtypes = [C.c_uint16]
        for n in ("pqto_dist", "pqto_dist_host"):
            getattr(L, n).restype = C.c_float
            getattr(L, n).argtypes = [C.c_float] * 4
        L.pqto_project.restype = C.c_float
        L.pqto_project.argtypes = [C.c_float] * 3
        for n in ("pqto_project_d", "pqto_project_d_host"):
            getattr(L, n).restype = C.c_float
            getattr(L, n).argtypes = [C.c_float] * 3 + [C.POINTER(C.c_float)]
        L.pqto_seg_dist.restype = C.c_float
        L.pqto_seg_dist.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.pqto_dist_seq.restype = C.c_uint32
        L.pqto_dist_seq.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
        L.pqto_query_knn.restype = C.c_int
        L.pqto_line_adc.restype = C.c_float
        _lib = L
    return _lib


def ref_host():
    """The reference's own triangle.cuh / bitonicSort.cuh compiled for the host
    (None when oracle/_ref was never built, e.g. no reference tree)."""
    if not os.path.exists(REF_HOST_PATH):
        return None
    L = C.CDLL(REF_HOST_PATH)
    L.ref_toUShort.restype = C.c_uint16
    L.ref_toUShort.argtypes = [C.c_float]
    L.ref_toFloat.restype = C.c_float
    L.ref_toFloat.argtypes = [C.c_uint16]
    L.ref_dist.restype = C.c_float
    L.ref_dist.argtypes = [C.c_float] * 4
    L.ref_project.restype = C.c_float
    L.ref_project.argtypes = [C.c_float] * 3
    L.ref_project_d.restype = C.c_float
    L.ref_project_d.argtypes = [C.c_float] * 3 + [C.POINTER(C.c_float)]
    L.ref_equal.restype = C.c_int
    L.ref_equal.argtypes = [C.c_float, C.c_float]
    L.ref_bitonic.restype = C.c_int
    L.ref_bitonic.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint32]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def default_params(dim, p, c1, c2, line_parts, **over):
    prm = Params()
    lib().pqto_default_params(C.byref(prm), dim, p, c1, c2, line_parts)
    for k, v in over.items():
        setattr(prm, k, v)
    return prm


def pow2ceil(x):
    return int(lib().pqto_pow2ceil(x))


def bitonic(val, idx):
    val = _f32(val).copy()
    idx = _u32(idx).copy()
    lib().pqto_bitonic(_p(val), _p(idx), C.c_uint32(val.size))
    return val, idx


def scan(v, inclusive):
    v = _u32(v).copy()
    lib().pqto_scan(_p(v), C.c_uint32(v.size), C.c_int(1 if inclusive else 0))
    return v


def seg_dist(q, c):
    q, c = _f32(q), _f32(c)
    return float(lib().pqto_seg_dist(_p(q), _p(c), q.size))


def dist_seq(max_cluster, p):
    seq = np.zeros(NUM_DISTSEQ, np.uint32)
    nv = C.c_uint32(0)
    m = lib().pqto_dist_seq(max_cluster, p, _p(seq), C.byref(nv))
    return seq, int(m), int(nv.value)


def cb_dist(prm, cb1):
    cb1 = _f32(cb1)
    out = np.zeros((prm.c1, prm.c1, prm.line_parts), np.float32)
    lib().pqto_cb_dist(C.byref(prm), _p(cb1), _p(out))
    return out


def line_adc(prm, lut, cbd, code):
    lut, cbd, code = _f32(lut), _f32(cbd), _u32(code)
    return float(lib().pqto_line_adc(C.byref(prm), _p(lut), _p(cbd), _p(code)))


def query_knn(prm, cb1, cb2, prefix, counts, db_idx, lines, Q, k, stages=False, nthreads=0):
    """Returns (dist[QN][k], idx[QN][k]) and, with stages=True, a dict of the
    per-query intermediates of SURVEY.md App. B."""
    cb1, cb2, Q = _f32(cb1), _f32(cb2), _f32(Q)
    prefix, counts, db_idx, lines = _u32(prefix), _u32(counts), _u32(db_idx), _u32(lines)
    QN = Q.shape[0]
    out_d = np.zeros((QN, k), np.float32)
    out_i = np.zeros((QN, k), np.uint32)
    st = None
    bufs = {}
    if stages:
        n = prm.k1 * prm.c2
        mv = pow2ceil(k)
        bufs = dict(
            assign=np.zeros((QN, prm.k1, prm.p), np.uint32),
            lut=np.zeros((QN, prm.line_parts, prm.c1), np.float32),
            assign_val=np.zeros((QN, prm.p, n), np.float32),
            assign_idx=np.zeros((QN, prm.p, n), np.uint32),
            bins=np.zeros((QN, prm.max_bins), np.uint32),
            n_bins=np.zeros(QN, np.uint32),
            select_idx=np.zeros((QN, mv), np.uint32),
            n_vec=np.zeros(QN, np.uint32))
        st = Stages(**{k_: a.ctypes.data for k_, a in bufs.items()})
    rc = lib().pqto_query_knn(C.byref(prm), _p(cb1), _p(cb2), _p(prefix), _p(counts), _p(db_idx),
                              _p(lines), _p(Q), C.c_uint32(QN), C.c_uint32(k), _p(out_d),
                              _p(out_i), C.byref(st) if st is not None else None,
                              C.c_int(nthreads))
    if rc != 0:
        raise ValueError("pqto_query_knn: unsupported shape (rc=%d)" % rc)
    return (out_d, out_i, bufs) if stages else (out_d, out_i)


NUM_ANISO_DIR = 10


def dist_seq_2d(max_cluster=512):
    seq = np.zeros(NUM_ANISO_DIR * NUM_DISTSEQ, np.uint32)
    lib().pqto_dist_seq_2d(C.c_uint32(max_cluster), _p(seq))
    return seq.reshape(NUM_ANISO_DIR, NUM_DISTSEQ)


def slope_idx(val0, val1, n):
    val0, val1 = _f32(val0), _f32(val1)
    amb = C.c_int(0)
    lib().pqto_slope_idx.restype = C.c_uint32
    si = lib().pqto_slope_idx(_p(val0), _p(val1), C.c_uint32(n), C.byref(amb))
    return int(si), bool(amb.value)


def big_params(dim, p, c1, c2, line_parts, **over):
    """literals of queryBIGKNNRerank2 (:8604-8639) and getBIGBins2D (:3725-3727)"""
    kw = dict(k1=16, max_bins=64 * 8192, max_trials=2560, bin_threads=1024)
    kw.update(over)
    return default_params(dim, p, c1, c2, line_parts, **kw)


def query_big_knn_rerank2(prm, cb1, cb2, prefix, counts, db_idx, lines, Q, k, nthreads=0):
    """Returns (dist, idx, info) with info = dict(n_bins, n_vec, ambiguous)."""
    cb1, cb2, Q = _f32(cb1), _f32(cb2), _f32(Q)
    prefix, counts, db_idx, lines = _u32(prefix), _u32(counts), _u32(db_idx), _u32(lines)
    QN = Q.shape[0]
    out_d = np.zeros((QN, k), np.float32)
    out_i = np.zeros((QN, k), np.uint32)
    nb = np.zeros(QN, np.uint32)
    nv = np.zeros(QN, np.uint32)
    amb = np.zeros(QN, np.int32)
    rc = lib().pqto_query_big_knn_rerank2(
        C.byref(prm), _p(cb1), _p(cb2), _p(prefix), _p(counts), _p(db_idx), _p(lines), _p(Q),
        C.c_uint32(QN), C.c_uint32(k), _p(out_d), _p(out_i), _p(nb), _p(nv), _p(amb),
        C.c_int(nthreads))
    if rc != 0:
        raise ValueError("pqto_query_big_knn_rerank2: unsupported shape (rc=%d)" % rc)
    return out_d, out_i, dict(n_bins=nb, n_vec=nv, ambiguous=(amb & 1).astype(bool),
                              ran_off_table=(amb & 2).astype(bool))


def assign_bins(prm, cb1, cb2, X, k1=16, nthreads=0):
    cb1, cb2, X = _f32(cb1), _f32(cb2), _f32(X)
    out = np.zeros(X.shape[0], np.uint32)
    lib().pqto_assign_bins(C.byref(prm), _p(cb1), _p(cb2), _p(X), C.c_uint32(X.shape[0]),
                           C.c_uint32(k1), _p(out), C.c_int(nthreads))
    return out


def build_lists(bin_of, hash_size):
    bin_of = _u32(bin_of)
    counts = np.zeros(hash_size, np.uint32)
    prefix = np.zeros(hash_size, np.uint32)
    db_idx = np.zeros(bin_of.size, np.uint32)
    lib().pqto_build_lists(_p(bin_of), C.c_uint32(bin_of.size), C.c_uint32(hash_size),
                           _p(counts), _p(prefix), _p(db_idx))
    return prefix, counts, db_idx


def line_encode(prm, cb1, cbd, X, nthreads=0):
    cb1, cbd, X = _f32(cb1), _f32(cbd), _f32(X)
    out = np.zeros((X.shape[0], prm.line_parts), np.uint32)
    lib().pqto_line_encode(C.byref(prm), _p(cb1), _p(cbd), _p(X), C.c_uint32(X.shape[0]),
                           _p(out), C.c_int(nthreads))
    return out


def brute_force_1nn(X, Q, nthreads=0):
    X, Q = _f32(X), _f32(Q)
    out = np.zeros(Q.shape[0], np.uint32)
    lib().pqto_brute_force_1nn(_p(X), C.c_uint32(X.shape[0]), _p(Q), C.c_uint32(Q.shape[0]),
                               C.c_uint32(X.shape[1]), _p(out), C.c_int(nthreads))
    return out


def build_index(prm, cb1, cb2, X, k1_build=16, nthreads=0):
    """tool_createdb's intent: bins, inverted lists, line codes."""
    bin_of = assign_bins(prm, cb1, cb2, X, k1_build, nthreads)
    prefix, counts, db_idx = build_lists(bin_of, prm.hash_size)
    cbd = cb_dist(prm, cb1)
    lines = line_encode(prm, cb1, cbd, X, nthreads)
    return dict(bin_of=bin_of, prefix=prefix, counts=counts, db_idx=db_idx, lines=lines,
                cb_dist=cbd)
