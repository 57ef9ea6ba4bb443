"""The drop-in boundary without a GPU: libpqt_b200.so loads, exports every symbol
include/pqt_b200.h declares, and fails loudly (no CPU fallback) when no device exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import conftest
import pqt_b200
from pqt_b200 import formats

HEADER = os.path.join(conftest.ROOT, "include", "pqt_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pqt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    pqt_b200.build()
    L = C.CDLL(pqt_b200.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 28
    for n in names:
        assert hasattr(L, n), "libpqt_b200.so lacks %s" % n
    assert set(names) == set(pqt_b200.EXPORTS)
    assert L.pqt_abi_version() == 1


def test_struct_layouts_match_the_header():
    assert C.sizeof(pqt_b200.Params) == 16 * 4
    assert C.sizeof(pqt_b200.Stats) == 4 * 8 + 5 * 8 + 8 + 7 * 8
    prm = pqt_b200.Params()
    pqt_b200.lib().pqt_default_params(C.byref(prm))
    # the reference's literals
    assert (prm.k1, prm.max_bins, prm.max_trials, prm.bin_threads, prm.max_vec_per_bin,
            prm.hash_size, prm.k1_build, prm.max_vec) == (8, 4096, 16, 1024, 2800, 400000000, 16, 0)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pqt_b200.PqtError):
        pqt_b200.PerturbationProTree(128, 4)
    h = C.c_void_p()
    assert pqt_b200.lib().pqt_create(128, 4, 2, 0, C.byref(h)) != 0  # p2 != p
    assert pqt_b200.lib().pqt_last_error(None) == b"null handle"


def test_product_does_not_touch_the_oracle():
    # the product path must never include / import / link / load anything under oracle/
    pkg = os.path.join(conftest.ROOT, "product-quantization-tree_b200")
    bad = re.compile(r"#\s*include[^\n]*oracle|import\s+pqt_oracle|from\s+pqt_oracle|"
                     r"libpqt_oracle|-lpqt_oracle|CDLL\([^)]*oracle|pqto_[a-z_]+\s*\(")
    seen = 0
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hh", ".hpp", ".h")) or f == "Makefile":
                seen += 1
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not bad.search(txt), os.path.join(dp, f)
    assert seen >= 8


def test_mem_and_ppqt_formats(tmp_path):
    rng = np.random.default_rng(0)
    x = rng.integers(0, 256, (37, 128)).astype(np.uint8)
    p = str(tmp_path / "base.umem")
    formats.write_mem(p, x)
    raw = open(p, "rb").read()
    assert raw[:7] == b"37\n128\n" and raw[7:20] == b"\0" * 13 and len(raw) == 20 + 37 * 128
    assert formats.read_mem_header(p) == (37, 128)
    assert np.array_equal(formats.read_mem(p, np.uint8), x)
    assert np.array_equal(formats.read_umem_as_float(p, 5, 3), x[3:8].astype(np.float32))
    gt = rng.integers(0, 1000, (5, 100)).astype(np.int32)
    formats.write_mem(str(tmp_path / "gt.imem"), gt)
    assert np.array_equal(formats.read_mem(str(tmp_path / "gt.imem"), np.int32), gt)
    cb1 = rng.normal(size=(16, 128)).astype(np.float32)
    cb2 = rng.normal(size=(4, 16, 8, 32)).astype(np.float32)
    formats.write_ppqt(str(tmp_path / "t.ppqt"), 128, 4, cb1, cb2)
    t = formats.read_ppqt(str(tmp_path / "t.ppqt"))
    assert (t["dim"], t["p"], t["p2"], t["c1"], t["c2"], t["nDBs"]) == (128, 4, 4, 16, 8, 1)
    assert np.array_equal(t["cb1"], cb1) and np.array_equal(t["cb2"], cb2)
    assert formats.base_name("tmp", 128, 4, 32, 32) == "tmp_128_4_32_32"
    assert formats.index_paths("tmp_128_4_32_32", 16)["lines"] == "tmp_128_4_32_32_16.lines"
