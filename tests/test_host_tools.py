"""The C++ host side: tool_createdb / tool_query (same flags and file names as the
reference's tools) over the C ABI."""
import os
import subprocess

import numpy as np
import pytest

import conftest
import pqt_oracle as po
from pqt_b200 import formats, synth

PKG = os.path.join(conftest.ROOT, "product-quantization-tree_b200")
TOOL_QUERY = os.path.join(PKG, "tool_query")
TOOL_CREATEDB = os.path.join(PKG, "tool_createdb")


def _build():
    subprocess.check_call(["make", "-C", PKG, "--no-print-directory"], stdout=subprocess.DEVNULL)


def test_tools_build_and_keep_the_reference_flags():
    _build()
    for tool, extra in ((TOOL_QUERY, ["queryset"]), (TOOL_CREATEDB, [])):
        out = subprocess.run([tool, "--help"], capture_output=True, text=True)
        assert out.returncode == 0
        # tool_query.cpp:26-36 / tool_createdb.cpp:26-35
        for flag in ["device", "c1", "c2", "p", "dim", "lineparts", "chunksize", "hashsize",
                     "basename", "dataset"] + extra:
            assert "--" + flag in out.stdout
    bad = subprocess.run([TOOL_QUERY, "--nonsense", "1"], capture_output=True, text=True)
    assert bad.returncode != 0 and "unknown flag" in bad.stderr


def test_tools_fail_loudly_without_a_device(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    _build()
    x = synth.db_vectors(0, 64, 128, 16)
    formats.write_mem(str(tmp_path / "base.umem"), x)
    r = subprocess.run([TOOL_CREATEDB, "--dataset", str(tmp_path / "base.umem"), "--c1", "16",
                        "--c2", "8", "--p", "4", "--chunksize", "64", "--hashsize", "4099",
                        "--basename", str(tmp_path / "t")], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr  # no CPU fallback


@pytest.mark.gpu
def test_createdb_then_query_end_to_end(tmp_path):
    _build()
    N, QN, dim, p, c1, c2, LP, hs, k = 8000, 40, 128, 4, 16, 8, 16, 65537, 256
    mu = synth.centres(64, dim)
    X = synth.db_vectors(0, N, dim, 64, mu=mu)
    Q, _ = synth.query_vectors(QN, N, dim, 64, mu=mu)
    formats.write_mem(str(tmp_path / "base.umem"), X)
    formats.write_mem(str(tmp_path / "query.umem"), Q)
    base = str(tmp_path / "tmp")
    common = ["--c1", str(c1), "--c2", str(c2), "--p", str(p), "--dim", str(dim), "--lineparts",
              str(LP), "--hashsize", str(hs), "--chunksize", "3000", "--basename", base,
              "--dataset", str(tmp_path / "base.umem")]
    r = subprocess.run([TOOL_CREATEDB] + common + ["--compact", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    pre = formats.base_name(base, dim, p, c1, c2)
    paths = formats.index_paths(pre, LP)
    tree = formats.read_ppqt(paths["ppqt"])
    assert (tree["dim"], tree["p"], tree["c1"], tree["c2"], tree["nDBs"]) == (dim, p, c1, c2, 1)
    prefix = np.fromfile(paths["prefix"], np.uint32)
    counts = np.fromfile(paths["count"], np.uint32)
    db_idx = np.fromfile(paths["dbIdx"], np.uint32)
    lines = np.fromfile(paths["lines"], np.uint32).reshape(N, LP)
    assert prefix.size == hs and counts.size == hs and db_idx.size == N
    # the files hold what the oracle's builder produces from the same codebooks
    prm = po.default_params(dim, p, c1, c2, LP, hash_size=hs)
    ref = po.build_index(prm, tree["cb1"], tree["cb2"], X.astype(np.float32), k1_build=16)
    assert np.array_equal(counts, ref["counts"]) and np.array_equal(prefix, ref["prefix"])
    assert np.array_equal(db_idx, ref["db_idx"]) and np.array_equal(lines, ref["lines"])

    out = str(tmp_path / "res")
    r = subprocess.run([TOOL_QUERY] + common + ["--queryset", str(tmp_path / "query.umem"), "--k",
                                                str(k), "--out", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "queries/s" in r.stdout
    idx = formats.read_mem(out + ".idx.imem", np.uint32)
    dist = formats.read_mem(out + ".dist.fmem", np.float32)
    d0, i0 = po.query_knn(prm, tree["cb1"], tree["cb2"], prefix, counts, db_idx, lines,
                          Q.astype(np.float32), k)
    assert np.array_equal(idx, i0) and np.array_equal(dist, d0)

    # the same query from the compact index file (one file, the resident layout as it is)
    assert os.path.getsize("%s_%d.pqtx" % (pre, LP)) < 4 * (2 * hs + N + N * LP)  # smaller than the four files
    out2 = str(tmp_path / "res2")
    r = subprocess.run([TOOL_QUERY] + common + ["--queryset", str(tmp_path / "query.umem"), "--k", str(k),
                                                "--out", out2, "--compact", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert np.array_equal(formats.read_mem(out2 + ".idx.imem", np.uint32), i0)
    assert np.array_equal(formats.read_mem(out2 + ".dist.fmem", np.float32), d0)
