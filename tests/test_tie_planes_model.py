"""Model check (CPU, numpy) of the bit-plane network simulation used by
product-quantization-tree_b200/csrc/tie_resolve.cuh.

The CUDA code re-creates the order that the reference's bitonic network
(pqt/bitonicSort.cuh:16-44: pairs (i, i^j), ascending iff (i & k) == 0, swap on strict > / <)
gives to candidates with bit-equal distances by running the network on three bit planes
(Lower / Higher than the tie value, label of the Equal positions) instead of on the data.
This test restates the plane update rules word for word in numpy and checks them against a
direct simulation of the network on (value, id) pairs with injected ties."""
import numpy as np

import conftest  # noqa: F401  (sys.path)
import pqt_oracle as po

U = np.uint32
MASKS = {1: 0x55555555, 2: 0x33333333, 4: 0x0F0F0F0F, 8: 0x00FF00FF, 16: 0x0000FFFF}


def reference_network(val, idx):
    n = len(val)
    val, idx = val.copy(), idx.copy()
    k = 2
    while k <= n:
        j = k >> 1
        while j > 0:
            i = np.arange(n)
            ixj = i ^ j
            m = ixj > i
            a, b = i[m], ixj[m]
            asc = (a & k) == 0
            sw = np.where(asc, val[a] > val[b], val[a] < val[b])
            a_s, b_s = a[sw], b[sw]
            val[a_s], val[b_s] = val[b_s].copy(), val[a_s].copy()
            idx[a_s], idx[b_s] = idx[b_s].copy(), idx[a_s].copy()
            j >>= 1
        k <<= 1
    return val, idx


def planes_network(L, H, B, n):
    """tie_substage_inword / tie_substage_xword applied over all sub-stages."""
    nW = n >> 5
    L, H, B = L.copy(), H.copy(), B.copy()
    x = np.arange(nW, dtype=np.uint32)
    k = 2
    while k <= n:
        j = k >> 1
        while j > 0:
            if j < 32:
                M = U(MASKS[j])
                if k < 32:
                    D = np.full(nW, (~U(MASKS[k])) & M, dtype=np.uint32)
                else:
                    D = np.where((x & (k >> 5)) != 0, M, U(0)).astype(np.uint32)
                aL, bL = L & M, (L >> U(j)) & M
                aH, bH = H & M, (H >> U(j)) & M
                aB, bB = B & M, (B >> U(j)) & M
                gt = (aH & ~bH) | (~aL & bL)
                lt = (bH & ~aH) | (~bL & aL)
                sw = (gt & ~D) | (lt & D)
                dL, dH, dB = (aL ^ bL) & sw, (aH ^ bH) & sw, (aB ^ bB) & sw
                L = L ^ (dL | (dL << U(j)))
                H = H ^ (dH | (dH << U(j)))
                B = B ^ (dB | (dB << U(j)))
            else:
                wd = j >> 5
                oL, oH, oB = L[x ^ wd], H[x ^ wd], B[x ^ wd]
                lo = (x & wd) == 0
                desc = (x & (k >> 5)) != 0
                aL, bL = np.where(lo, L, oL), np.where(lo, oL, L)
                aH, bH = np.where(lo, H, oH), np.where(lo, oH, H)
                gt = (aH & ~bH) | (~aL & bL)
                lt = (bH & ~aH) | (~bL & aL)
                sw = np.where(desc, lt, gt)
                L = L ^ ((L ^ oL) & sw)
                H = H ^ ((H ^ oH) & sw)
                B = B ^ ((B ^ oB) & sw)
            j >>= 1
        k <<= 1
    return L, H, B


def _bits(mask):
    return np.packbits(mask.reshape(-1, 32)[:, ::-1], axis=1).view(">u4").reshape(-1).astype(np.uint32)


def test_bit_plane_network_reproduces_the_tie_order_of_the_reference_network():
    rng = np.random.default_rng(1)
    ID_A, ID_B = 100000, 100001
    for trial in range(24):
        n = int(rng.choice([128, 1024, 4096]))
        nv = int(rng.integers(n // 2, n + 1))
        val = np.full(n, 1e7, np.float32)  # pads, as in rerankKernelFast (:5333)
        val[:nv] = rng.permutation(nv).astype(np.float32) * 3.0 + 1
        idx = np.arange(n)
        m = int(rng.integers(2, 9))  # tie group: m candidates, two vectors (with duplicates)
        pos = rng.choice(nv, m, replace=False)
        v = val[pos[0]]
        val[pos] = v
        lab = rng.integers(0, 2, m)
        lab[0], lab[1] = 0, 1
        idx[pos] = np.where(lab == 0, ID_A, ID_B)
        _, ref_idx = reference_network(val, idx)
        # the numpy network above is the oracle's network (itself pinned to the reference's
        # bitonicSort.cuh in tests/test_oracle_known_answers.py)
        _, oi = po.bitonic(val, idx.astype(np.uint32))
        assert np.array_equal(oi, ref_idx.astype(np.uint32))
        L, H, B = planes_network(_bits(val < v), _bits(val > v), _bits((val == v) & (idx == ID_B)), n)
        eq = ~(L | H)
        r = int((val < v).sum())  # the group's output slots are r .. r+m-1
        assert sum(bin(int(w)).count("1") for w in eq) == m
        for e in range(r, r + m):
            w, b = e >> 5, e & 31
            assert (eq[w] >> U(b)) & 1
            assert (ID_B if (B[w] >> U(b)) & 1 else ID_A) == ref_idx[e]
