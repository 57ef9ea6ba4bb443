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


def _popc(a):
    return np.array([bin(int(w)).count("1") for w in a], dtype=np.int64)


def _low(n):
    n = np.clip(n, 0, 32)
    return np.where(n >= 32, 0xFFFFFFFF, (1 << n) - 1).astype(np.uint32)


def fast_forward(L, H, B, n, max_words=32):
    """The closed form tie_resolve.cuh starts from: the planes after stage kk0 = 32 * bw0, bw0
    the largest block size (in words, up to max_words = one warp) at which no aligned block
    holds two Equal positions.  Returns (L, H, B, first stage to simulate)."""
    nW = n >> 5
    x = np.arange(nW, dtype=np.int64)
    cL, cH, cE, cB = _popc(L), _popc(H), _popc(~(L | H)), _popc(B)
    bw0 = 0
    bw = 1
    while bw <= max_words and bw <= nW:
        if (cE.reshape(-1, bw).sum(1) > 1).any():
            break
        bw0 = bw
        bw <<= 1
    if bw0 == 0:
        return L, H, B, 2
    blk = x // bw0
    nL = cL.reshape(-1, bw0).sum(1)[blk]
    nH = cH.reshape(-1, bw0).sum(1)[blk]
    nE = cE.reshape(-1, bw0).sum(1)[blk]
    lab = cB.reshape(-1, bw0).sum(1)[blk] != 0
    o = (x & (bw0 - 1)) << 5
    asc = (x & bw0) == 0
    L2 = np.where(asc, _low(nL - o), ~_low(nH + nE - o)).astype(np.uint32)
    H2 = np.where(asc, ~_low(nL + nE - o), _low(nH - o)).astype(np.uint32)
    B2 = np.where(lab, ~(L2 | H2), 0).astype(np.uint32)
    return L2, H2, B2, bw0 << 6


def planes_network(L, H, B, n, k_first=2):
    """tie_substage_inword / tie_substage_xword applied over all sub-stages from stage k_first."""
    nW = n >> 5
    L, H, B = L.copy(), H.copy(), B.copy()
    x = np.arange(nW, dtype=np.uint32)
    k = k_first
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
    for trial in range(36):
        n = int(rng.choice([64, 128, 1024, 4096]))
        nv = int(rng.integers(n // 2, n + 1))
        val = np.full(n, 1e7, np.float32)  # pads, as in rerankKernelFast (:5333)
        val[:nv] = rng.permutation(nv).astype(np.float32) * 3.0 + 1
        idx = np.arange(n)
        m = int(rng.integers(2, 9))  # tie group: m candidates, two vectors (with duplicates)
        if trial % 3 == 0:    # anywhere
            pos = rng.choice(nv, m, replace=False)
        elif trial % 3 == 1:  # neighbours in candidate order (vectors of one bin): the common case
            p0 = int(rng.integers(0, nv - m + 1))
            pos = np.arange(p0, p0 + m)
        else:                 # two far-apart elements: most stages are fast-forwarded
            m = 2
            pos = np.array([int(rng.integers(0, nv // 2)), int(rng.integers(nv // 2, nv))])
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
        L0, H0, B0 = _bits(val < v), _bits(val > v), _bits((val == v) & (idx == ID_B))
        L, H, B = planes_network(L0, H0, B0, n)
        # the fast-forwarded start (closed form after the last stage whose blocks hold at most
        # one Equal position) ends in the same planes
        Lf, Hf, Bf, k_first = fast_forward(L0, H0, B0, n)
        Lf, Hf, Bf = planes_network(Lf, Hf, Bf, n, k_first)
        assert np.array_equal(Lf, L) and np.array_equal(Hf, H)
        assert np.array_equal(Bf & ~(Lf | Hf), B & ~(L | H))
        eq = ~(L | H)
        r = int((val < v).sum())  # the group's output slots are r .. r+m-1
        assert sum(bin(int(w)).count("1") for w in eq) == m
        for e in range(r, r + m):
            w, b = e >> 5, e & 31
            assert (eq[w] >> U(b)) & 1
            assert (ID_B if (B[w] >> U(b)) & 1 else ID_A) == ref_idx[e]
