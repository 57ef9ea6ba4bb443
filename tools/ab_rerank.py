#!/usr/bin/env python
"""A/B timing of the fused scan+rank kernel variants on one resident index.
usage: python tools/ab_rerank.py [bench.py flags]   (env switches are set per run)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "product-quantization-tree_b200", "python"))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
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
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ref = None
    for mode, tpb, dd, bd in (("split", "512", "1", "0"), ("split", "512", "1", "1"), ("fused", "512", "1", "1")):
        if True:
            os.environ["PQT_SCAN_MODE"] = mode
            os.environ["PQT_BINS_DEDUPE"] = bd
            os.environ["PQT_RERANK_TPB"] = tpb
            os.environ["PQT_RERANK_DEDUPE"] = dd
            for _ in range(3):
                t.queryKNN(Qd, a.qn, a.k, oi, od)
            t.profile(True)
            t.reset_stats()
            for _ in range(a.steps):
                flush.zero_()
                t.queryKNN(Qd, a.qn, a.k, oi, od)
            torch.cuda.synchronize()
            st = t.stats()
            t.profile(False)
            cur = (oi.clone(), od.clone())
            same = None if ref is None else bool(torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1]))
            ref = ref or cur
            print("bins_dedupe %s " % bd, end="")
            print("%s tpb %s dedupe %s: tables %.3f bins %.3f scan %.3f rank %.3f ms/step  same_as_first=%s" % (
                mode, tpb, dd, st.ms_tables / a.steps, st.ms_bins / a.steps, st.ms_scan / a.steps,
                st.ms_sort / a.steps, same), flush=True)
    t.close()


if __name__ == "__main__":
    main()
