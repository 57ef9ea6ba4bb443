"""Bench / test plumbing around the synthetic database (not part of the product):

* the generator of pqt_b200/synth.py as a CUDA kernel (libpqt_synth.so) writing uint8 chunks
  straight into torch tensors, so that a 1-B-vector base set never exists anywhere as a whole;
* the cluster-centre table and exact 1-NN ground truth in torch;
* tool_synthdb (native, over libpqt_b200.so): writes the reference's index files
  (.prefix / .count / .dbIdx / .lines) of the synthetic database for the CPU arms and for the
  loader path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpqt_synth.so")
TOOL_PATH = os.path.join(HERE, "tool_synthdb")
M32 = 0xFFFFFFFF

_lib = None


def build():
    subprocess.check_call(["make", "-C", HERE, "--no-print-directory"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("tools/synthdb/libpqt_synth.so is not built (make -C tools/synthdb)")
        L = C.CDLL(LIB_PATH)
        L.pqts_db_u8.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p,
                                 C.c_uint32, C.c_uint32, C.c_void_p]
        _lib = L
    return _lib


def centres_u8(n_clusters, dim, seed, device):
    """pqt_b200.synth.centres as a torch uint8 tensor on `device` (float64 log like numpy)."""
    import torch

    def fmix32(x):
        x = x & M32
        x = x ^ (x >> 16)
        x = (x * 0x85EBCA6B) & M32
        x = x ^ (x >> 13)
        x = (x * 0xC2B2AE35) & M32
        return x ^ (x >> 16)

    out = torch.empty((n_clusters, dim), dtype=torch.uint8, device=device)
    d = torch.arange(dim, dtype=torch.int64, device=device)[None, :]
    step = 1 << 18
    for s in range(0, n_clusters, step):
        g = torch.arange(s, min(n_clusters, s + step), dtype=torch.int64, device=device)[:, None]
        hh = fmix32(fmix32(seed ^ ((0xC3A5C85C + g) & M32)) + d * 0x9E3779B9)
        u = (hh.to(torch.float64) + 1.0) / 4294967297.0
        mu = torch.clamp(torch.round(-28.0 * torch.log(u)), max=218.0)
        out[s:s + g.shape[0]] = mu.to(torch.uint8)
    return out


def db_u8(out, i0, n, mu, seed, stream=None):
    """rows i0 .. i0+n-1 of the synthetic database into the CUDA uint8 tensor `out` [>= n][dim]"""
    import torch
    assert out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and mu.is_cuda
    dim = out.shape[1]
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    rc = lib().pqts_db_u8(out.data_ptr(), i0, n, dim, mu.data_ptr(), mu.shape[0], seed, st)
    if rc != 0:
        raise RuntimeError("pqts_db_u8 failed: cuda error %d" % rc)
    return out[:n]


def queries_u8(nq, n_db, mu, seed, qseed):
    """pqt_b200.synth.query_vectors over a uint8 centre table on any device: perturbed copies of
    random database vectors.  Returns (uint8 [nq][dim], source ids int64)."""
    import torch
    from pqt_b200 import synth_torch as st
    dev = mu.device
    dim = mu.shape[1]
    j = torch.arange(nq, dtype=torch.int64, device=dev)
    src = st.fmix32(st.fmix32(j) ^ qseed) % n_db
    g = st.cluster_of(src, mu.shape[0], seed)
    x = (mu[g].to(torch.int64) + st._noise(seed, src, dim, 10)).clamp_(0, 255)
    x = (x + st._noise(qseed, j, dim, 11)).clamp_(0, 255)
    return x.to(torch.uint8), src


def exact_1nn(Q8, n, mu, seed, chunk=1 << 18, i_lo=0, i_hi=None):
    """Exact nearest neighbour (squared L2 over uint8 coordinates, ties -> lowest id) of every
    query among database vectors i_lo .. i_hi-1, regenerating the base set chunk by chunk.
    Returns (score float32 [QN], id int64 [QN]) on the queries' device; score = |x|^2 - 2 q.x
    orders the candidates of one query.

    One augmented GEMM per chunk gives the whole score: rows [2x, a0, a1, a2] against
    [-q, 1, 256, 65536] with |x|^2 = a0 + 256 a1 + 65536 a2.  Every operand is an integer below
    2^11 in magnitude or a power of two and every partial sum stays below 2^24, so the product is
    exact in TF32 inputs / fp32 accumulation: tensor-core speed, exact integers."""
    import torch
    dev = Q8.device
    i_hi = n if i_hi is None else i_hi
    QN, dim = Q8.shape
    K = (dim + 3 + 7) // 8 * 8
    qa = torch.zeros((QN, K), dtype=torch.float32, device=dev)
    qa[:, :dim] = -Q8.to(torch.float32)
    qa[:, dim] = 1.0
    qa[:, dim + 1] = 256.0
    qa[:, dim + 2] = 65536.0
    best = torch.full((QN,), float("inf"), dtype=torch.float32, device=dev)
    arg = torch.zeros((QN,), dtype=torch.int64, device=dev)
    m_max = min(chunk, max(1, i_hi - i_lo))
    buf = torch.empty((m_max, dim), dtype=torch.uint8, device=dev)
    xa = torch.zeros((m_max, K), dtype=torch.float32, device=dev)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    for s in range(i_lo, i_hi, chunk):
        m = min(chunk, i_hi - s)
        X = db_u8(buf, s, m, mu, seed)
        xi = X.to(torch.int32)
        x2 = (xi * xi).sum(1)
        xa[:m, :dim] = (2 * xi).to(torch.float32)
        xa[:m, dim] = (x2 & 255).to(torch.float32)
        xa[:m, dim + 1] = ((x2 >> 8) & 255).to(torch.float32)
        xa[:m, dim + 2] = (x2 >> 16).to(torch.float32)
        sc = qa @ xa[:m].t()
        v, i = sc.min(1)
        upd = v < best
        best = torch.where(upd, v, best)
        arg = torch.where(upd, i.to(torch.int64) + s, arg)
    torch.backends.cuda.matmul.allow_tf32 = prev
    return best, arg


def index_files(basename, dim, p, c1, c2, lineparts):
    pre = "%s_%d_%d_%d_%d" % (basename, dim, p, c1, c2)
    return dict(pre=pre, ppqt=pre + ".ppqt", prefix=pre + ".prefix", count=pre + ".count",
                dbIdx=pre + ".dbIdx", lines="%s_%d.lines" % (pre, lineparts))


def files_complete(paths, n, hashsize, lineparts):
    want = {"prefix": hashsize * 4, "count": hashsize * 4, "dbIdx": n * 4, "lines": n * lineparts * 4}
    return all(os.path.exists(paths[k]) and os.path.getsize(paths[k]) == v for k, v in want.items())


def run_tool(basename, n, dim, p, c1, c2, lineparts, hashsize, clusters, seed, mu_path, device=0,
             chunksize=10000000, log=None):
    """tool_synthdb as a child process (the calling process never maps libpqt_b200.so)."""
    if not os.path.exists(TOOL_PATH):
        raise ImportError("tools/synthdb/tool_synthdb is not built (make -C tools/synthdb)")
    cmd = [TOOL_PATH, "--n", str(n), "--dim", str(dim), "--p", str(p), "--c1", str(c1), "--c2", str(c2),
           "--lineparts", str(lineparts), "--hashsize", str(hashsize), "--clusters", str(clusters),
           "--seed", str(seed), "--mu", mu_path, "--basename", basename, "--device", str(device),
           "--chunksize", str(chunksize)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if log is not None:
        log.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("tool_synthdb failed (%d): %s" % (r.returncode, (r.stderr or r.stdout)[-2000:]))
    return r.stdout
