#!/usr/bin/env python
"""Where does the CUDA path differ from the oracle on the bench workload?
usage: python tools/parity_probe.py [bench.py flags] (uses --cpu-sample queries, default 4096)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "product-quantization-tree_b200", "python"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    a = bench.parse()
    if a.clusters <= 0:
        a.clusters = max(4096, a.n // 256)
    X8, Q8, src, cb1, cb2 = bench.build_inputs(a, "cuda:0")
    t, _ = bench.build_index_gpu(a, X8, cb1, cb2, 0)
    del X8
    Qd = Q8.to(torch.float32).contiguous()
    oi = torch.empty((a.qn, a.k), dtype=torch.int32, device="cuda")
    od = torch.empty((a.qn, a.k), dtype=torch.float32, device="cuda")
    t.queryKNN(Qd, a.qn, a.k, oi, od)
    torch.cuda.synchronize()
    gi = oi.cpu().numpy().view(np.uint32)
    gd = od.cpu().numpy()
    po = bench.oracle_handles()
    prefix, counts, db_idx = t.getDB()
    lines = t.getLine()
    host_index = dict(prefix=prefix, counts=counts, db_idx=db_idx, lines=lines)
    S = a.cpu_sample or 4096
    _, d0, i0 = bench.cpu_baseline(a, po, host_index, cb1, cb2, Qd.cpu().numpy(), S, os.cpu_count())
    bad = [q for q in range(S) if not (np.array_equal(i0[q], gi[q]) and np.array_equal(d0[q], gd[q]))]
    print("queries compared", S, "differing", len(bad))
    for q in bad[:12]:
        nv = int((i0[q] != 0xFFFFFFFF).sum())
        dd = np.nonzero(d0[q] != gd[q])[0]
        di = np.nonzero(i0[q] != gi[q])[0]
        print("q=%d nv=%d  dist diffs %d (first %s)  idx diffs %d (first %s)" %
              (q, nv, dd.size, dd[:4].tolist(), di.size, di[:6].tolist()))
        for e in di[:3]:
            lo, hi = max(0, e - 2), min(a.k, e + 3)
            print("   slot %d: oracle d=%s i=%s" % (e, d0[q][lo:hi].tolist(), i0[q][lo:hi].tolist()))
            print("            cuda   d=%s i=%s" % (gd[q][lo:hi].tolist(), gi[q][lo:hi].tolist()))
    t.close()


if __name__ == "__main__":
    main()
