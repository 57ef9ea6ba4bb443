"""Debug helper (GPU box): run the reference's kernels twice on the test case of
tests/test_ref_gpu.py and report run-to-run determinism and agreement with the oracle."""
import os, sys, subprocess, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); import conftest
import pqt_oracle as po
from pqt_b200 import formats, synth
tmp = '/tmp/refdbg'; os.makedirs(tmp, exist_ok=True)
HASH = 400000000
N, QN, dim, p, c1, c2, LP, k1, k, KB = 20000, 48, 128, 4, 16, 8, 16, 8, 1024, 16
mu = synth.centres(256, dim)
X = synth.db_vectors(0, N, dim, 256, mu=mu).astype(np.float32)
Q = synth.query_vectors(QN, N, dim, 256, mu=mu)[0].astype(np.float32)
cb1, cb2 = synth.train_tree(X[:5000], p, c1, c2, iters=6, seed=5)
ppqt = tmp + '/ref_128_4_16_8.ppqt'
formats.write_ppqt(ppqt, dim, p, cb1, cb2)
prm = po.default_params(dim, p, c1, c2, LP, hash_size=HASH)
index = po.build_index(prm, cb1, cb2, X, k1_build=16)
nz = np.nonzero(index["counts"])[0].astype(np.uint32)
np.savez(tmp + '/case.npz', X=X, Q=Q, dim=dim, p=p, c1=c1, c2=c2, LP=LP, k1=k1, k=k, hash_size=HASH, k_big=KB,
         nz_bins=nz, nz_counts=index["counts"][nz], db_idx=index["db_idx"], lines=index["lines"])
outs = []
for rep in range(2):
    out = tmp + '/out%d.npz' % rep
    subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'ref_gpu_runner.py'), tmp + '/case.npz', out, ppqt],
                   capture_output=True, text=True, timeout=600)
    outs.append(dict(np.load(out + '.big.npz')))
bp = po.big_params(dim, p, c1, c2, LP, hash_size=HASH)
d0, i0, info = po.query_big_knn_rerank2(bp, cb1, cb2, index["prefix"], index["counts"], index["db_idx"], index["lines"], Q, KB)
print("ref run-to-run n_bins equal:", np.array_equal(outs[0]['big_n_bins'], outs[1]['big_n_bins']))
print("ref0 n_bins", outs[0]['big_n_bins'][:24])
print("ref1 n_bins", outs[1]['big_n_bins'][:24])
print("orcl n_bins", info['n_bins'][:24])
print("ran_off   ", info['ran_off_table'][:24].astype(int))
print("ambiguous ", info['ambiguous'][:24].astype(int))
bad = [q for q in range(QN) if outs[0]['big_n_bins'][q] != info['n_bins'][q] and not info['ran_off_table'][q]]
print("mismatching queries (not ran_off):", bad)
for q in bad[:3]:
    print(q, "ref bins", outs[0]['big_bins'][q][:20], "nvec", info['n_vec'][q])
