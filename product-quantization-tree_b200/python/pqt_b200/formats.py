"""The reference's on-disk formats, numpy side.

.umem/.imem/.fmem: ASCII header "<N>\\n<D>\\n", zero padding up to byte 20, payload
from byte 20 (writer convert/filehelper.hpp:251-278, reader utils/filereader.hpp:56-69).
.ppqt: ASCII "dim p p2 c1 c2 nDBs" lines, one skipped byte, float cb1[c1][dim],
float cb2[p][c1][c2][dim/p] (pqt/PerturbationProTree.cu:60-220).
.prefix/.count/.dbIdx/.lines: raw little-endian arrays (tool_createdb.cpp:116-138).
"""
import numpy as np

HEADER_BYTES = 20


def write_mem(path, arr):
    arr = np.ascontiguousarray(arr)
    assert arr.ndim == 2
    hdr = ("%d\n%d\n" % arr.shape).encode()
    assert len(hdr) <= HEADER_BYTES, "header does not fit the fixed 20-byte prefix"
    with open(path, "wb") as f:
        f.write(hdr + b"\0" * (HEADER_BYTES - len(hdr)))
        f.write(arr.tobytes())


def read_mem_header(path):
    with open(path, "rb") as f:
        toks = f.read(HEADER_BYTES).split()
    return int(toks[0]), int(toks[1])


def read_mem(path, dtype, num=None, offset=0):
    n, d = read_mem_header(path)
    num = n - offset if num is None else num
    dt = np.dtype(dtype)
    with open(path, "rb") as f:
        f.seek(HEADER_BYTES + offset * d * dt.itemsize)
        data = np.fromfile(f, dtype=dt, count=num * d)
    return data.reshape(num, d)


def read_umem_as_float(path, num=None, offset=0):
    """FileReader<float>::data: uint8 payload widened to float (utils/filereader.hpp:34-49)."""
    return read_mem(path, np.uint8, num, offset).astype(np.float32)


def write_ppqt(path, dim, p, cb1, cb2):
    cb1 = np.ascontiguousarray(cb1, np.float32)
    cb2 = np.ascontiguousarray(cb2, np.float32)
    c1 = cb1.shape[0]
    c2 = cb2.shape[2]
    with open(path, "wb") as f:
        f.write(("%d\n%d\n%d\n%d\n%d\n%d\n" % (dim, p, p, c1, c2, 1)).encode())
        f.write(cb1.tobytes())
        f.write(cb2.tobytes())


def read_ppqt(path):
    with open(path, "rb") as f:
        raw = f.read()
    vals, pos = [], 0
    for _ in range(6):
        end = raw.index(b"\n", pos)
        vals.append(int(raw[pos:end]))
        pos = end + 1
    dim, p, p2, c1, c2, ndb = vals
    n1 = ndb * c1 * dim
    n2 = ndb * c1 * c2 * dim
    cb1 = np.frombuffer(raw, np.float32, n1, pos).reshape(c1, dim).copy()
    cb2 = np.frombuffer(raw, np.float32, n2, pos + 4 * n1).reshape(p, c1, c2, dim // p).copy()
    return dict(dim=dim, p=p, p2=p2, c1=c1, c2=c2, nDBs=ndb, cb1=cb1, cb2=cb2)


def base_name(basename, dim, p, c1, c2):
    """tool_query.cpp:77-78"""
    return "%s_%d_%d_%d_%d" % (basename, dim, p, c1, c2)


def index_paths(pre, lineparts):
    """tool_query.cpp:93,105-108 / tool_createdb.cpp:81-84"""
    return dict(ppqt=pre + ".ppqt", lines="%s_%d.lines" % (pre, lineparts),
                prefix=pre + ".prefix", count=pre + ".count", dbIdx=pre + ".dbIdx")
