import os, sys, subprocess, numpy as np
sys.path.insert(0,'/root/repo/tests'); import conftest
import pqt_oracle as po
from pqt_b200 import formats, synth
tmp='/tmp/refdbg'; os.makedirs(tmp, exist_ok=True)
HASH=400000000
N, QN, dim, p, c1, c2, LP, k1, k = 20000, 48, 128, 4, 16, 8, 16, 8, 1024
mu = synth.centres(256, dim)
X = synth.db_vectors(0, N, dim, 256, mu=mu).astype(np.float32)
Q = synth.query_vectors(QN, N, dim, 256, mu=mu)[0].astype(np.float32)
cb1, cb2 = synth.train_tree(X[:5000], p, c1, c2, iters=6, seed=5)
ppqt = tmp+'/ref_128_4_16_8.ppqt'
formats.write_ppqt(ppqt, dim, p, cb1, cb2)
prm = po.default_params(dim, p, c1, c2, LP, hash_size=HASH)
index = po.build_index(prm, cb1, cb2, X, k1_build=16)
nz = np.nonzero(index["counts"])[0].astype(np.uint32)
np.savez(tmp+'/case.npz', X=X, Q=Q, dim=dim, p=p, c1=c1, c2=c2, LP=LP, k1=k1, k=k, hash_size=HASH, nz_bins=nz, nz_counts=index["counts"][nz], db_idx=index["db_idx"], lines=index["lines"])
outs=[]
for rep in range(2):
    out=tmp+'/out%d.npz'%rep
    r=subprocess.run([sys.executable,'/root/repo/tests/ref_gpu_runner.py',tmp+'/case.npz',out,ppqt],capture_output=True,text=True,timeout=600)
    outs.append(dict(np.load(out)))
d0,i0,st0 = po.query_knn(prm, cb1, cb2, index["prefix"], index["counts"], index["db_idx"], index["lines"], Q, k, stages=True)
np.savez_compressed('/root/repo/gpurun_out/refdbg.npz', ref_idx0=outs[0]['idx'], ref_dist0=outs[0]['dist'], ref_idx1=outs[1]['idx'], ref_dist1=outs[1]['dist'], d0=d0, i0=i0, n_vec=st0['n_vec'], select_idx=st0['select_idx'], n_bins=st0['n_bins'])
print("run-to-run equal:", np.array_equal(outs[0]['dist'],outs[1]['dist']), np.array_equal(outs[0]['idx'],outs[1]['idx']))
print("equal to oracle:", (outs[0]['dist']==d0).mean())
