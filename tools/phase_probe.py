#!/usr/bin/env python
"""Per-query phase timing of the fused scan+rank kernel (debug stage rerank_phases).
usage: python tools/phase_probe.py [bench.py flags, e.g. --n 1000000]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "product-quantization-tree_b200", "python"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    sys.argv = [sys.argv[0]] + sys.argv[1:]
    a = bench.parse()
    import pqt_b200
    inp = bench.build_inputs(a, "cuda:0")
    t = pqt_b200.PerturbationProTree(a.dim, a.p, a.p, 0)
    t.set_params(hash_size=a.hashsize, k1_build=min(16, a.c1))
    t.setTree(inp["cb1"], inp["cb2"])
    bench.build_index_chunked(a, t, inp, 0, 1, "cuda:0")
    Qd = inp["Q8"].to(torch.float32).contiguous()
    oi = torch.empty((a.qn, a.k), dtype=torch.int32, device="cuda")
    od = torch.empty((a.qn, a.k), dtype=torch.float32, device="cuda")
    for _ in range(2):
        t.queryKNN(Qd, a.qn, a.k, oi, od)
    t.debug(True)
    t.queryKNN(Qd, a.qn, a.k, oi, od)
    torch.cuda.synchronize()
    ph = t.stage("rerank_phases", (a.qn, 8), np.uint64).astype(np.int64)
    t.debug(False)
    split = bool((ph[:, 7] < 0).any())  # bit 63: stamps of the ranking-only kernel (split pipeline)
    ph[:, 7] &= (1 << 63) - 1
    nv = (ph[:, 7] & 0x3FFF) >> 1
    tie_search = ((ph[:, 7] >> 14) & 0x3FFFF) << 4
    tie_groups = (ph[:, 7] >> 48) & 15
    fast = ph[:, 7] & 1
    sm = (ph[:, 7] >> 32) & 0xFFFF
    names = ["lut_wait", "scan", "sort+emit", "repair", "ties", "network/tail"]
    if split:
        # rank2_kernel: slot 1 = end of the sort proper; phases: load, sort, emit, repair, ties, tail
        names = ["load", "sort", "emit", "repair", "ties", "network/tail"]
        s1 = np.where(ph[:, 1] > 0, ph[:, 1], ph[:, 2])
        d = np.stack([ph[:, 2] - ph[:, 0], s1 - ph[:, 2],
                      np.where(ph[:, 4] > 0, ph[:, 4] - s1, 0),
                      np.where(ph[:, 3] > 0, ph[:, 3] - ph[:, 4], 0),
                      np.where(ph[:, 5] > 0, ph[:, 5] - ph[:, 3], 0),
                      ph[:, 6] - np.where(ph[:, 5] > 0, ph[:, 5], ph[:, 2])], 1)
    else:
      d = np.stack([ph[:, 1] - ph[:, 0], ph[:, 2] - ph[:, 1],
                  np.where(ph[:, 4] > 0, ph[:, 4] - ph[:, 2], 0),
                  np.where(ph[:, 3] > 0, ph[:, 3] - ph[:, 4], 0),
                  np.where(ph[:, 5] > 0, ph[:, 5] - ph[:, 3], 0),
                  ph[:, 6] - np.where(ph[:, 5] > 0, ph[:, 5], ph[:, 2])], 1)
    tot = ph[:, 6] - ph[:, 0]
    print("queries", a.qn, "fast", int(fast.sum()), "mean nv", nv.mean())
    print("mean cycles per query: total %.0f" % tot.mean())
    for i, n in enumerate(names):
        print("  %-22s mean %8.0f  p99 %8.0f  max %9d" % (n, d[:, i].mean(), np.percentile(d[:, i], 99), d[:, i].max()))
    if split:
        ties = d[:, 4]
        print("  tie resolver: search for groups mean %.0f cycles; groups per query: %s" %
              (tie_search.mean(), " ".join("%d:%d" % (g, int((tie_groups == g).sum())) for g in range(9))))
        for g in range(0, 9):
            m = tie_groups == g
            if m.any():
                print("    %d groups: %5d queries, ties phase mean %.0f (search %.0f)" % (g, m.sum(), ties[m].mean(), tie_search[m].mean()))
    for lo, hi in [(0, 512), (512, 1024), (1024, 2048), (2048, 4097)]:
        m = (nv >= lo) & (nv < hi)
        if m.any():
            print("  nv in [%d,%d): %5d queries, total mean %.0f | " % (lo, hi, m.sum(), tot[m].mean()) +
                  " ".join("%s %.0f" % (n, d[m, i].mean()) for i, n in enumerate(names)))
    worst = np.argsort(-tot)[:8]
    for q in worst:
        print("  worst q=%d sm=%d nv=%d fast=%d total=%d phases=%s" % (q, sm[q], nv[q], fast[q], tot[q], d[q].tolist()))
    per_sm = np.zeros(int(sm.max()) + 1)
    np.add.at(per_sm, sm, tot)
    print("per-SM busy cycles (sum over both groups): min %.0f mean %.0f max %.0f" % (per_sm.min(), per_sm.mean(), per_sm.max()))
    t.close()


if __name__ == "__main__":
    main()
