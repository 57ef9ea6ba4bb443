#!/usr/bin/env python
"""Times the REFERENCE's own CUDA path on the B200: pqt/*.cu recompiled unmodified for sm_100a
(oracle/_ref/libpqt_ref_gpu.so, see oracle/Makefile) running pqt::PerturbationProTree::queryKNN
(pqt/PerturbationProTree.cu:8179-8323) with its mallocs, host-side prepareDistSequence, device
printfs and result copy -- the "Maxwell kernels on B200" number of BASELINE.md 2.3.
Workload = BASELINE configs[1] (1M x 128, lineparts 16, HASH_SIZE 4e8 is compiled into the
reference).  Informational: the reference's rerankKernelFast is racy on this GPU,
so only the timing and recall@1 are reported, not parity.
usage: python tools/ref_gpu_bench.py [--runs 3] [--qn 10000]      (prints one JSON line)"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "product-quantization-tree_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tools", "synthdb"))
import numpy as np  # noqa: E402

LIB = os.path.join(ROOT, "oracle", "_ref", "libpqt_ref_gpu.so")


def child(paths_json, qn, k, runs):
    paths = json.loads(paths_json)
    L = C.CDLL(LIB)
    L.refgpu_create.restype = C.c_void_p
    L.refgpu_hash_size.restype = C.c_uint32
    vp = C.c_void_p
    hs = int(L.refgpu_hash_size())
    prefix = np.fromfile(paths["prefix"], np.uint32)
    counts = np.fromfile(paths["count"], np.uint32)
    dbidx = np.fromfile(paths["dbIdx"], np.uint32)
    lines = np.fromfile(paths["lines"], np.uint32)
    assert prefix.size == hs
    N = dbidx.size
    LP = lines.size // N
    Q = np.load(paths["queries"])
    h = vp(L.refgpu_create(128, 4))
    L.refgpu_read_tree(h, paths["ppqt"].encode())
    L.refgpu_set_db(h, N, vp(prefix.ctypes.data), vp(counts.ctypes.data), vp(dbidx.ctypes.data))
    L.refgpu_set_lines(h, vp(lines.ctypes.data), N, LP)
    idx = np.zeros((qn, k), np.uint32)
    dist = np.zeros((qn, k), np.float32)
    times = []
    for _ in range(runs + 1):  # first call = warm-up
        t0 = time.perf_counter()
        L.refgpu_query_knn(h, vp(Q.ctypes.data), qn, k, vp(idx.ctypes.data), vp(dist.ctypes.data))
        times.append(time.perf_counter() - t0)
    np.save(paths["out"], idx[:, 0].copy())
    sys.stderr.write("REFGPU_TIMES " + json.dumps(times[1:]) + "\n")


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=3)
    ap.add_argument("--qn", type=int, default=10000)
    ap.add_argument("--k", type=int, default=4096)
    ap.add_argument("--child", default="")
    a = ap.parse_args()
    if a.child:
        child(a.child, a.qn, a.k, a.runs)
        return
    if not os.path.exists(LIB):
        print(json.dumps({"impl": "reference-gpu", "unavailable": "oracle/_ref/libpqt_ref_gpu.so is not built"}))
        return
    import torch
    import bench
    import synthdb
    sys.argv = [sys.argv[0], "--workload", "c2", "--qn", str(a.qn), "--k", str(a.k)]
    b = bench.parse()
    inp = bench.build_inputs(b, "cuda:0")
    paths, _, _ = bench.ensure_index_files(b, inp, 0)
    gt = synthdb.exact_1nn(inp["Q8"], b.n, inp["mu"], bench.DB_SEED)[1].cpu().numpy().astype(np.uint32)
    qpath = os.path.join(paths["dir"], "queries_f32.npy")
    np.save(qpath, inp["Q8"].to(torch.float32).cpu().numpy())
    p = dict(paths, queries=qpath, out=os.path.join(paths["dir"], "refgpu_top1.npy"))
    del inp
    torch.cuda.empty_cache()
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", json.dumps(p), "--qn", str(a.qn),
                        "--k", str(a.k), "--runs", str(a.runs)], stdout=subprocess.DEVNULL,
                       stderr=subprocess.PIPE, text=True, timeout=900)
    times = None
    for line in r.stderr.splitlines():
        if line.startswith("REFGPU_TIMES "):
            times = json.loads(line[len("REFGPU_TIMES "):])
    if r.returncode != 0 or not times:
        print(json.dumps({"impl": "reference-gpu", "unavailable": "reference run failed: " + r.stderr[-300:]}))
        return
    top1 = np.load(p["out"])
    best = min(times)
    print(json.dumps({
        "impl": "reference GPU path: pqt/*.cu recompiled unmodified for sm_100a, PerturbationProTree::queryKNN "
                "(host buffers in, std::vector out, as tool_query calls it)",
        "workload": bench.workload_name(b), "runs": a.runs, "seconds_per_call": times,
        "queries_per_s_best": a.qn / best, "queries_per_s_mean": a.qn / (sum(times) / len(times)),
        "recall_at_1": float((top1 == gt).mean()),
        "note": "stock path incl. 8 cudaMalloc/cudaFree, host prepareDistSequence, device printf and the "
                "328 MB result copy per call; rerankKernelFast is racy on this GPU (distances past the "
                "first 64 candidates vary run to run)"}))


if __name__ == "__main__":
    main()
