"""Pins the oracle to the reference's own known answers (run.cu:9-115,
pqt/bitonicSort.cuh:213-252) and, where oracle/_ref exists, to the reference's own
headers compiled for the host (oracle/ref_shim.cpp)."""
import ctypes as C

import numpy as np
import pytest

import pqt_oracle as po


def test_pow2ceil_is_the_reference_log2():
    # pqt/helper.hh:27-37 returns the next power of two
    for x, want in [(1, 1), (2, 2), (3, 4), (16, 16), (17, 32), (256, 256), (4096, 4096),
                    (4097, 8192), (2800, 4096)]:
        assert po.pow2ceil(x) == want


# run.cu:33-104: six (a2, b2, c2) -> (lambda, d2) cases, tolerance 1e-5 (equal(), triangle.cuh:112)
TRIANGLES = [
    (1.0, 2.0, 1.0, 1.0, 1.0),
    (2.0, 2.0, 4.0, 0.5, 1.0),
    (2.0, 2.0, 2.0, 0.5, 1.5),
    (2.0, 5.0, 9.0, 0.666666666, 1.0),
    (2.0, 5.0, 1.0, 2.0, 1.0),
    (5.0, 2.0, 1.0, -1.0, 1.0),
]


@pytest.mark.parametrize("a2,b2,c2,lam,d2", TRIANGLES)
def test_run_cu_triangle_cases(a2, b2, c2, lam, d2):
    L = po.lib()
    d = C.c_float()
    l = L.pqto_project_d(a2, b2, c2, C.byref(d))
    assert abs(l - lam) < 1e-5
    assert abs(d.value - d2) < 1e-5
    dh = C.c_float()
    assert L.pqto_project_d_host(a2, b2, c2, C.byref(dh)) == l and abs(dh.value - d2) < 1e-5
    # "d2 == dist(lambda)" identity of run.cu, for the device (FMA) and host forms
    assert abs(L.pqto_dist(a2, b2, c2, l) - d.value) < 1e-5
    assert abs(L.pqto_dist_host(a2, b2, c2, l) - d.value) < 1e-5
    assert abs(L.pqto_project(a2, b2, c2) - lam) < 1e-5


def test_lambda_quantiser_range_and_step():
    # run.cu:106-112 + pqt/triangle.cuh:6-18: [-4, 4) in steps of 8/65536, clamped
    L = po.lib()
    assert L.pqto_to_ushort(-4.0) == 0
    assert L.pqto_to_ushort(-10.0) == 0
    assert L.pqto_to_ushort(4.0) == 65535
    assert L.pqto_to_ushort(10.0) == 65535
    assert L.pqto_to_ushort(0.0) == 32768
    assert L.pqto_to_ushort(float("nan")) == 0
    for i in range(-100, 100):
        f = np.float32(i / 10.0)
        back = L.pqto_to_float(L.pqto_to_ushort(f))
        if -4.0 <= f < 4.0:
            assert 0.0 <= f - back < 8.0 / 65536 + 1e-7  # truncation
        elif f >= 4.0:
            assert back == np.float32(65535 * (8.0 / 65536) - 4.0)
        else:
            assert back == -4.0
    s = np.arange(65536, dtype=np.uint32)
    vals = np.array([L.pqto_to_float(int(x)) for x in s[::257]], np.float32)
    assert np.all(np.diff(vals) > 0)
    for x in s[::257]:
        assert L.pqto_to_ushort(L.pqto_to_float(int(x))) == x  # exact round trip


@pytest.mark.parametrize("n", [1024, 2048, 4096])
def test_sort_test_large_ramp(n):
    # sortTestLarge (pqt/bitonicSort.cuh:213-232): val = N - tid, idx = tid -> idx[tid] == N-tid-1
    val = (n - np.arange(n)).astype(np.float32)
    idx = np.arange(n, dtype=np.uint32)
    v, i = po.bitonic(val, idx)
    assert np.array_equal(i, (n - np.arange(n) - 1).astype(np.uint32))
    assert np.all(np.diff(v) > 0)


@pytest.mark.parametrize("n", [1024, 2048, 4096])
def test_scan_test_large_ones(n):
    # scanTestLarge (:234-252): exclusive scan of ones == tid; ProTree::testScan
    # (pqt/ProTree.cu:1755-1816) checks inclusive + exclusive against the host
    ones = np.ones(n, np.uint32)
    assert np.array_equal(po.scan(ones, False), np.arange(n, dtype=np.uint32))
    assert np.array_equal(po.scan(ones, True), np.arange(1, n + 1, dtype=np.uint32))
    rng = np.random.default_rng(n)
    v = rng.integers(0, 2801, n).astype(np.uint32)
    assert np.array_equal(po.scan(v, True), np.cumsum(v, dtype=np.uint64).astype(np.uint32))
    assert np.array_equal(po.scan(v, False),
                          (np.cumsum(v, dtype=np.uint64) - v).astype(np.uint32))


def test_kbest_order_against_std_sort():
    # ProQuantization::testKBestAssignment (pqt/ProQuantization.cu:747-819): device k-best
    # order vs host std::sort (20 % slack there; exact here because values are distinct)
    rng = np.random.default_rng(5)
    for n in (16, 32, 256):
        val = rng.permutation(n * 4)[:n].astype(np.float32)
        v, i = po.bitonic(val, np.arange(n, dtype=np.uint32))
        order = np.argsort(val, kind="stable")
        assert np.array_equal(i, order.astype(np.uint32))
        assert np.array_equal(v, val[order])


# ---- the reference's own headers compiled for the host --------------------------------

ref = po.ref_host()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref not built (no reference tree)")


@needs_ref
def test_triangle_bitwise_against_reference_header():
    rng = np.random.default_rng(11)
    L = po.lib()
    for _ in range(4000):
        a2, b2, c2 = (np.float32(x) for x in rng.uniform(0.0, 5e4, 3))
        lam = np.float32(rng.uniform(-4.5, 4.5))
        # host compilation of triangle.cuh is uncontracted -> compare with the host form
        assert L.pqto_dist_host(a2, b2, c2, lam) == ref.ref_dist(a2, b2, c2, lam)
        assert L.pqto_project(a2, b2, c2) == ref.ref_project(a2, b2, c2)
        d0, d1 = C.c_float(), C.c_float()
        l0 = L.pqto_project_d_host(a2, b2, c2, C.byref(d0))
        l1 = ref.ref_project_d(a2, b2, c2, C.byref(d1))
        assert l0 == l1 and d0.value == d1.value
        # device form: the final mul+sub is one FFMA there
        d2 = C.c_float()
        assert L.pqto_project_d(a2, b2, c2, C.byref(d2)) == l1
        assert abs(d2.value - d1.value) <= 1e-6 * max(1.0, abs(float(b2)), abs(float(c2)) * l1 * l1)
        # the device form differs from the host form by FMA rounding only
        dev = L.pqto_dist(a2, b2, c2, lam)
        host = ref.ref_dist(a2, b2, c2, lam)
        assert abs(dev - host) <= 1e-5 * max(1.0, abs(host))
    for f in np.linspace(-6, 6, 4001, dtype=np.float32):
        assert L.pqto_to_ushort(f) == ref.ref_toUShort(f)
    for s in range(0, 65536, 7):
        assert L.pqto_to_float(s) == ref.ref_toFloat(s)


@needs_ref
def test_run_cu_cases_on_reference_header():
    for a2, b2, c2, lam, d2 in TRIANGLES:
        d = C.c_float()
        l = ref.ref_project_d(a2, b2, c2, C.byref(d))
        assert ref.ref_equal(l, lam) and ref.ref_equal(d.value, d2)
        assert ref.ref_equal(ref.ref_dist(a2, b2, c2, l), d.value)


@needs_ref
@pytest.mark.parametrize("n,large,threads", [(32, 0, 0), (256, 0, 0), (1024, 1, 16),
                                             (4096, 1, 8)])
def test_bitonic_network_against_reference_header(n, large, threads):
    # the reference's bitonic3 / bitonicLarge run by real host threads; inputs with many
    # exact ties so that the network's (unstable) tie order is pinned too
    rng = np.random.default_rng(n + large)
    for trial in range(3):
        val = rng.integers(0, n // 4, n).astype(np.float32)
        if trial == 2:
            val[n // 2:] = 1e7  # padding pattern of rerankKernelFast
        idx = rng.permutation(n).astype(np.uint32)
        v_ref, i_ref = val.copy(), idx.copy()
        rc = ref.ref_bitonic(v_ref.ctypes.data, i_ref.ctypes.data, n, large, threads)
        assert rc == 0
        v, i = po.bitonic(val, idx)
        assert np.array_equal(v, v_ref)
        assert np.array_equal(i, i_ref)
        assert np.all(np.diff(v) >= 0)


@needs_ref
def test_sort_test_large_on_reference_header():
    n = 2048
    val = (n - np.arange(n)).astype(np.float32)
    idx = np.arange(n, dtype=np.uint32)
    assert ref.ref_bitonic(val.ctypes.data, idx.ctypes.data, n, 1, 8) == 0
    assert np.array_equal(idx, (n - np.arange(n) - 1).astype(np.uint32))
