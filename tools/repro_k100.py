"""repro: split pipeline, k < nVec (max_vec extension) on a dense index"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "product-quantization-tree_b200", "python"))
import torch, numpy as np
import bench
sys.argv = [sys.argv[0], "--n", os.environ.get("REPRO_N", "20000000"), "--qn", os.environ.get("REPRO_QN", "256")]
a = bench.parse()
import pqt_b200
inp = bench.build_inputs(a, "cuda:0")
t = pqt_b200.PerturbationProTree(a.dim, a.p, a.p, 0)
t.set_params(hash_size=a.hashsize, k1_build=16)
t.setTree(inp["cb1"], inp["cb2"])
bench.build_index_chunked(a, t, inp, 0, 1, "cuda:0")
Qd = inp["Q8"].to(torch.float32).contiguous()
for mode in ("fused", "split"):
    os.environ["PQT_SCAN_MODE"] = mode
    t.set_params(max_vec=0)
    oi = torch.empty((a.qn, 4096), dtype=torch.int32, device="cuda"); od = torch.empty((a.qn, 4096), dtype=torch.float32, device="cuda")
    t.queryKNN(Qd, a.qn, 4096, oi, od); torch.cuda.synchronize()
    print(mode, "k=4096 ok", flush=True)
    t.set_params(max_vec=4096)
    oi2 = torch.empty((a.qn, 100), dtype=torch.int32, device="cuda"); od2 = torch.empty((a.qn, 100), dtype=torch.float32, device="cuda")
    t.queryKNN(Qd, a.qn, 100, oi2, od2); torch.cuda.synchronize()
    print(mode, "k=100 device outputs ok", flush=True)
    pi = torch.empty((a.qn, 100), dtype=torch.int32).pin_memory(); pd = torch.empty((a.qn, 100), dtype=torch.float32).pin_memory()
    t.queryKNN(Qd.cpu().pin_memory(), a.qn, 100, pi, pd); torch.cuda.synchronize()
    print(mode, "k=100 host outputs ok:", bool(torch.equal(pi.cuda(), oi2) and torch.equal(pd.cuda(), od2)), flush=True)
    print(mode, "k=100 ok, equal to first 100 of k=4096:", bool(torch.equal(oi2, oi[:, :100].contiguous()) and torch.equal(od2, od[:, :100].contiguous())), flush=True)
t.close()
