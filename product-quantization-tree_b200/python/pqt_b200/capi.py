"""ctypes binding of libpqt_b200.so (include/pqt_b200.h).

No CPU fallback: if the library is missing, or no CUDA device is present,
constructing an index raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB_PATH = os.path.join(PKG_ROOT, "libpqt_b200.so")

PAD_IDX = 0xFFFFFFFF
PQT_OK = 0

EXPORTS = [
    "pqt_abi_version", "pqt_create", "pqt_destroy", "pqt_last_error", "pqt_default_params",
    "pqt_set_params", "pqt_get_params", "pqt_set_stream", "pqt_read_tree", "pqt_write_tree",
    "pqt_set_tree", "pqt_get_tree_shape", "pqt_get_tree", "pqt_set_db", "pqt_set_lines",
    "pqt_query_knn", "pqt_query_big_knn_rerank2", "pqt_build_kbest_db", "pqt_line_dist", "pqt_get_db", "pqt_get_lines",
    "pqt_get_db_size", "pqt_set_shard",
    "pqt_candidate_width", "pqt_shard_exchange_alloc", "pqt_shard_exchange_handle",
    "pqt_shard_exchange_open", "pqt_shard_exchange_set_peers", "pqt_shard_exchange_ptrs",
    "pqt_shard_dispatch", "pqt_shard_scan_p2p", "pqt_shard_rank", "pqt_profile_enable", "pqt_get_stats", "pqt_reset_stats",
    "pqt_debug_enable", "pqt_debug_stage",
    "pqt_assign_bins", "pqt_set_db_from_bins", "pqt_line_dist_begin", "pqt_line_dist_chunk",
    "pqt_line_dist_end", "pqt_get_codes_binorder", "pqt_save_index", "pqt_load_index",
]

X_F32, X_U8 = 0, 1

STAGES = dict(assign=0, lut=1, assign_val=2, assign_idx=3, bins=4, n_bins=5, select_idx=6,
              n_vec=7, cb_dist=8, dist_seq=9, dist_seq_2d=10, big_bins=11, big_n_bins=12, rerank_phases=13)


class Params(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "k1", "max_bins", "max_trials", "bin_threads", "max_vec_per_bin", "hash_size",
        "k1_build", "max_vec", "big_k1", "big_max_bins", "big_max_trials", "rank_mode")] + \
        [("reserved", C.c_uint32 * 4)]


class Stats(C.Structure):
    _fields_ = [("calls", C.c_uint64), ("queries", C.c_uint64), ("candidates", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("ms_tables", C.c_double),
                ("ms_bins", C.c_double), ("ms_scan", C.c_double), ("ms_sort", C.c_double),
                ("ms_total", C.c_double), ("scan_launches", C.c_uint64),
                ("exact_rank_queries", C.c_uint64), ("tie_resolved_queries", C.c_uint64),
                ("stream_scan_launches", C.c_uint64), ("reserved", C.c_uint64 * 4)]


class PqtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pqt_b200 error %d: %s" % (code, msg))
        self.code = code


def build(force=False):
    """make the in-tree shared library (nvcc, sm_100a)."""
    src = os.path.join(PKG_ROOT, "csrc")
    newest = max(os.path.getmtime(os.path.join(src, f)) for f in os.listdir(src))
    newest = max(newest, os.path.getmtime(os.path.join(PKG_ROOT, "..", "include", "pqt_b200.h")))
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        subprocess.check_call(["make", "-C", PKG_ROOT, "--no-print-directory"],
                              stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libpqt_b200.so is not built (run `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C product-quantization-tree_b200`); there is no CPU "
                "fallback")
        L = C.CDLL(LIB_PATH)
        L.pqt_last_error.restype = C.c_char_p
        L.pqt_last_error.argtypes = [C.c_void_p]
        L.pqt_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                 C.POINTER(C.c_void_p)]
        L.pqt_destroy.argtypes = [C.c_void_p]
        L.pqt_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.pqt_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
        L.pqt_get_params.argtypes = [C.c_void_p, C.POINTER(Params)]
        L.pqt_read_tree.argtypes = [C.c_void_p, C.c_char_p]
        L.pqt_write_tree.argtypes = [C.c_void_p, C.c_char_p]
        L.pqt_set_tree.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.pqt_get_tree_shape.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint32)] * 4
        L.pqt_get_tree.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pqt_set_db.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pqt_set_lines.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.pqt_query_knn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32,
                                    C.c_void_p, C.c_void_p, C.c_int]
        L.pqt_query_big_knn_rerank2.argtypes = L.pqt_query_knn.argtypes
        L.pqt_build_kbest_db.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32]
        L.pqt_line_dist.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32]
        L.pqt_assign_bins.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32,
                                      C.c_void_p, C.c_int]
        L.pqt_set_db_from_bins.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32]
        L.pqt_line_dist_begin.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.pqt_line_dist_chunk.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32,
                                          C.c_uint32, C.c_void_p]
        L.pqt_line_dist_end.argtypes = [C.c_void_p]
        L.pqt_get_codes_binorder.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        L.pqt_save_index.argtypes = [C.c_void_p, C.c_char_p]
        L.pqt_load_index.argtypes = [C.c_void_p, C.c_char_p]
        L.pqt_get_db.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pqt_get_lines.argtypes = [C.c_void_p, C.c_void_p]
        L.pqt_get_db_size.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.pqt_set_shard.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.pqt_shard_exchange_alloc.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.pqt_shard_exchange_handle.argtypes = [C.c_void_p, C.c_void_p]
        L.pqt_shard_exchange_open.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.pqt_shard_exchange_set_peers.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pqt_shard_exchange_ptrs.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                              C.POINTER(C.c_void_p)]
        L.pqt_shard_dispatch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32,
                                         C.c_uint32, C.c_uint32]
        L.pqt_shard_scan_p2p.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.pqt_shard_rank.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        L.pqt_candidate_width.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.pqt_profile_enable.argtypes = [C.c_void_p, C.c_int]
        L.pqt_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.pqt_reset_stats.argtypes = [C.c_void_p]
        L.pqt_debug_enable.argtypes = [C.c_void_p, C.c_int]
        L.pqt_debug_stage.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


def _ptr(x):
    """host numpy array, torch tensor (host or device), or raw int pointer -> (ptr, on_device)"""
    if x is None:
        return None, 0
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data, 0
    if isinstance(x, int):
        return x, 1
    # torch tensor (duck-typed; torch is plumbing, not a dependency of the binding)
    assert x.is_contiguous()
    return x.data_ptr(), 1 if x.is_cuda else 0


class PerturbationProTree:
    """Mirror of pqt::PerturbationProTree (pqt/PerturbationProTree.hh:28-235) over the C ABI:
    same method names and argument meaning as the reference class for the query path."""

    def __init__(self, dim, p, p2=None, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        rc = self._L.pqt_create(dim, p, p if p2 is None else p2, device, C.byref(self._h))
        if rc != PQT_OK:
            raise PqtError(rc, "pqt_create failed (no CUDA device, or p2 != p)")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.pqt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != PQT_OK:
            raise PqtError(rc, self._L.pqt_last_error(self._h).decode())

    # ---- parameters
    def params(self):
        prm = Params()
        self._chk(self._L.pqt_get_params(self._h, C.byref(prm)))
        return prm

    def set_params(self, **kw):
        prm = self.params()
        for k, v in kw.items():
            setattr(prm, k, v)
        self._chk(self._L.pqt_set_params(self._h, C.byref(prm)))

    def set_stream(self, cuda_stream_ptr):
        self._chk(self._L.pqt_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    # ---- tree
    def readTreeFromFile(self, name):
        self._chk(self._L.pqt_read_tree(self._h, name.encode()))

    def writeTreeToFile(self, name):
        self._chk(self._L.pqt_write_tree(self._h, name.encode()))

    def setTree(self, cb1, cb2):
        cb1 = np.ascontiguousarray(cb1, np.float32)
        cb2 = np.ascontiguousarray(cb2, np.float32)
        c1 = cb1.shape[0]
        c2 = cb2.shape[2]
        self._chk(self._L.pqt_set_tree(self._h, c1, c2, cb1.ctypes.data, cb2.ctypes.data))

    def shape(self):
        v = [C.c_uint32() for _ in range(4)]
        self._chk(self._L.pqt_get_tree_shape(self._h, *[C.byref(x) for x in v]))
        return tuple(int(x.value) for x in v)  # dim, p, c1, c2

    # ---- DB
    def setShard(self, rank, world):
        self._chk(self._L.pqt_set_shard(self._h, rank, world))

    def setDB(self, N, prefix, counts, dbIdx):
        prefix = np.ascontiguousarray(prefix, np.uint32)
        counts = np.ascontiguousarray(counts, np.uint32)
        dbIdx = np.ascontiguousarray(dbIdx, np.uint32)
        hs = self.params().hash_size
        assert prefix.size == hs and counts.size == hs and dbIdx.size == N
        self._chk(self._L.pqt_set_db(self._h, N, prefix.ctypes.data, counts.ctypes.data,
                                     dbIdx.ctypes.data))

    def setLines(self, lines, N, line_parts):
        lines = np.ascontiguousarray(lines, np.uint32)
        assert lines.size == N * line_parts
        self._chk(self._L.pqt_set_lines(self._h, lines.ctypes.data, N, line_parts))

    def buildKBestDB(self, A, N):
        ptr, dev = _ptr(A)
        self._chk(self._L.pqt_build_kbest_db(self._h, ptr, dev, N))

    def lineDist(self, DB, N, line_parts=16):
        ptr, dev = _ptr(DB)
        self._chk(self._L.pqt_line_dist(self._h, ptr, dev, N, line_parts))

    # ---- chunked build (test/test1B.cpp:783-871: the base set in 10-M-vector chunks)
    @staticmethod
    def _rows(X):
        """rows as float32 or uint8 (numpy host array or torch tensor) -> (ptr, kind, on_device)"""
        if isinstance(X, np.ndarray):
            if X.dtype != np.uint8:
                X = np.ascontiguousarray(X, np.float32)
            kind = X_U8 if X.dtype == np.uint8 else X_F32
            assert X.flags["C_CONTIGUOUS"]
            return X, X.ctypes.data, kind, 0
        kind = X_U8 if str(X.dtype) == "torch.uint8" else X_F32
        assert X.is_contiguous() and str(X.dtype) in ("torch.uint8", "torch.float32")
        return X, X.data_ptr(), kind, 1 if X.is_cuda else 0

    def assignBins(self, X, n, bin_out=None):
        """bins of n vectors; bin_out: torch CUDA int32 tensor / numpy uint32 array (returned)"""
        keep, xp, kind, dev = self._rows(X)
        if bin_out is None:
            bin_out = np.zeros(n, np.uint32)
        bp, bdev = _ptr(bin_out)
        self._chk(self._L.pqt_assign_bins(self._h, xp, kind, dev, n, bp, bdev))
        return bin_out

    def setDBFromBins(self, bin_of, N):
        bp, bdev = _ptr(bin_of)
        self._chk(self._L.pqt_set_db_from_bins(self._h, bp, bdev, N))

    def lineDistBegin(self, N, line_parts):
        self._chk(self._L.pqt_line_dist_begin(self._h, N, line_parts))

    def lineDistChunk(self, X, id0, n, lines_out=None):
        keep, xp, kind, dev = self._rows(X)
        lp = lines_out.ctypes.data if lines_out is not None else None
        self._chk(self._L.pqt_line_dist_chunk(self._h, xp, kind, dev, id0, n, lp))

    def lineDistEnd(self):
        self._chk(self._L.pqt_line_dist_end(self._h))

    def getCodesBinOrder(self, pos0=0, n=None, out=None):
        N, lp = self.dbSize()
        if n is None:
            n = N - pos0
        if out is None:
            out = np.zeros((n, lp), np.uint32)
        self._chk(self._L.pqt_get_codes_binorder(self._h, pos0, n, out.ctypes.data))
        return out

    def saveIndex(self, path):
        """the resident index (directory, dbIdx, bin-ordered codes of this handle's slice) as one file"""
        self._chk(self._L.pqt_save_index(self._h, os.fsencode(path)))

    def loadIndex(self, path):
        """restores a saveIndex file; the tree, hash_size and shard of the handle must match"""
        self._chk(self._L.pqt_load_index(self._h, os.fsencode(path)))

    def dbSize(self):
        n, lp = C.c_uint32(), C.c_uint32()
        self._chk(self._L.pqt_get_db_size(self._h, C.byref(n), C.byref(lp)))
        return int(n.value), int(lp.value)

    def getDB(self):
        N, _ = self.dbSize()
        hs = self.params().hash_size
        prefix = np.zeros(hs, np.uint32)
        counts = np.zeros(hs, np.uint32)
        db_idx = np.zeros(N, np.uint32)
        self._chk(self._L.pqt_get_db(self._h, prefix.ctypes.data, counts.ctypes.data,
                                     db_idx.ctypes.data))
        return prefix, counts, db_idx

    def getLine(self):
        N, lp = self.dbSize()
        lines = np.zeros((N, lp), np.uint32)
        self._chk(self._L.pqt_get_lines(self._h, lines.ctypes.data))
        return lines

    # ---- query
    def candidateWidth(self, k):
        mv = C.c_uint32()
        self._chk(self._L.pqt_candidate_width(self._h, k, C.byref(mv)))
        return int(mv.value)

    def queryKNN(self, Q, QN, k, out_idx=None, out_dist=None):
        """Q: numpy float32 [QN][dim] (host) or torch CUDA tensor (device, as in the
        reference).  Returns (resIdx, resDist) as [QN][k] arrays; pass torch CUDA
        tensors / pinned tensors as out_idx / out_dist to control placement."""
        if isinstance(Q, np.ndarray):
            Q = np.ascontiguousarray(Q, np.float32)
        qp, qdev = _ptr(Q)
        if out_idx is None:
            out_idx = np.zeros((QN, k), np.uint32)
            out_dist = np.zeros((QN, k), np.float32)
        ip, idev = _ptr(out_idx)
        dp, ddev = _ptr(out_dist)
        assert idev == ddev
        self._chk(self._L.pqt_query_knn(self._h, qp, qdev, QN, k, ip, dp, idev))
        return out_idx, out_dist

    def queryBIGKNNRerank2(self, Q, QN, k, out_idx=None, out_dist=None):
        """queryBIGKNNRerank2 (pqt/PerturbationProTree.hh:80): the 1-B variant; the line
        codes are the resident ones instead of the reference's pinned-host hLines."""
        if isinstance(Q, np.ndarray):
            Q = np.ascontiguousarray(Q, np.float32)
        qp, qdev = _ptr(Q)
        if out_idx is None:
            out_idx = np.zeros((QN, k), np.uint32)
            out_dist = np.zeros((QN, k), np.float32)
        ip, idev = _ptr(out_idx)
        dp, ddev = _ptr(out_dist)
        assert idev == ddev
        self._chk(self._L.pqt_query_big_knn_rerank2(self._h, qp, qdev, QN, k, ip, dp, idev))
        return out_idx, out_dist

    # ---- multi-GPU: dispatch + scan fused with the exchange over peer memory (include/pqt_b200.h)
    def shardExchangeAlloc(self, q_per_rank, max_vec):
        self._chk(self._L.pqt_shard_exchange_alloc(self._h, q_per_rank, max_vec))

    def shardExchangeHandle(self):
        buf = C.create_string_buffer(192)
        self._chk(self._L.pqt_shard_exchange_handle(self._h, buf))
        return buf.raw

    def shardExchangeOpen(self, handles):
        """handles: list of 192-byte IPC handle blobs, entry r from rank r"""
        blob = b"".join(handles)
        self._chk(self._L.pqt_shard_exchange_open(self._h, len(handles), blob))

    def shardExchangePtrs(self):
        v, i, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._chk(self._L.pqt_shard_exchange_ptrs(self._h, C.byref(v), C.byref(i), C.byref(c)))
        return v.value, i.value, c.value

    def shardExchangeSetPeers(self, val_ptrs, inbox_ptrs, cnt_ptrs):
        n = len(val_ptrs)
        va = (C.c_void_p * n)(*val_ptrs)
        ia = (C.c_void_p * n)(*inbox_ptrs)
        ca = (C.c_void_p * n)(*cnt_ptrs)
        self._chk(self._L.pqt_shard_exchange_set_peers(self._h, n, va, ia, ca))

    def shardDispatch(self, Q, QN, k, q_lo, q_hi):
        qp, qdev = _ptr(Q)
        self._chk(self._L.pqt_shard_dispatch(self._h, qp, qdev, QN, k, q_lo, q_hi))

    def shardScanP2P(self, QN, k):
        self._chk(self._L.pqt_shard_scan_p2p(self._h, QN, k))

    def shardRank(self, q_own, k, out_idx, out_dist):
        ip, idev = _ptr(out_idx)
        dp, _ = _ptr(out_dist)
        self._chk(self._L.pqt_shard_rank(self._h, q_own, k, ip, dp, idev))

    # ---- measurement / introspection
    def profile(self, on=True):
        self._chk(self._L.pqt_profile_enable(self._h, 1 if on else 0))

    def stats(self):
        st = Stats()
        self._chk(self._L.pqt_get_stats(self._h, C.byref(st)))
        return st

    def reset_stats(self):
        self._chk(self._L.pqt_reset_stats(self._h))

    def debug(self, on=True):
        self._chk(self._L.pqt_debug_enable(self._h, 1 if on else 0))

    def stage(self, name, shape, dtype):
        out = np.zeros(shape, dtype)
        self._chk(self._L.pqt_debug_stage(self._h, STAGES[name], out.ctypes.data, out.nbytes))
        return out
