"""Synthetic SIFT-shaped data and codebook training (inputs of the query path).

The generator is counter based and integer only, so any chunk of the database can
be regenerated bit-identically anywhere (numpy here, the same arithmetic in
csrc/synth_kernels.cuh on the GPU):

  h(x)      = murmur3 fmix32
  cluster g = h(seed ^ h(i)) mod G
  centre    mu_g[d] = min(218, round(-28 ln u)), u from h(seed_c, g, d)   (host table)
  x_i[d]    = clip(mu_g[d] + ((b0+b1+b2+b3 - 510) * 83 >> 10), 0, 255), b = bytes of
              h(h(seed + i) + d * 0x85EBCA77)        (sum of 4 uniform bytes ~ N(0, 12^2))
  query j   = clip(x_{r_j} + ((sum - 510) * 83 >> 11)) with r_j = h(seed_q ^ h(j)) mod N

Codebooks are inputs to parity (shared via .ppqt), so the k-means here only has to be
deterministic, not the reference's split-and-Lloyd (pqt/ProQuantization.cu:1047-1169).
"""
import numpy as np

DB_SEED = 20160627
QUERY_SEED = 424242
MASK = np.uint32(0xFFFFFFFF)


def fmix32(x):
    x = np.asarray(x, dtype=np.uint32).copy()
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x85EBCA6B)
    x ^= x >> np.uint32(13)
    x *= np.uint32(0xC2B2AE35)
    x ^= x >> np.uint32(16)
    return x


def centres(n_clusters, dim, seed=DB_SEED):
    g = np.arange(n_clusters, dtype=np.uint32)[:, None]
    d = np.arange(dim, dtype=np.uint32)[None, :]
    with np.errstate(over="ignore"):
        hh = fmix32(fmix32(np.uint32(seed) ^ np.uint32(0xC3A5C85C) + g) + d * np.uint32(0x9E3779B9))
    u = (hh.astype(np.float64) + 1.0) / 4294967297.0
    mu = np.minimum(218.0, np.rint(-28.0 * np.log(u)))
    return mu.astype(np.int32)


def _noise(seed, ids, dim, shift):
    with np.errstate(over="ignore"):
        base = fmix32(np.uint32(seed) + ids.astype(np.uint32))[:, None]
        d = np.arange(dim, dtype=np.uint32)[None, :]
        hh = fmix32(base + d * np.uint32(0x85EBCA77))
    s = ((hh & 0xFF) + ((hh >> 8) & 0xFF) + ((hh >> 16) & 0xFF) + (hh >> 24)).astype(np.int32)
    return ((s - 510) * 83) >> shift


def cluster_of(ids, n_clusters, seed=DB_SEED):
    with np.errstate(over="ignore"):
        return fmix32(np.uint32(seed) ^ fmix32(ids.astype(np.uint32))) % np.uint32(n_clusters)


def db_vectors(i0, n, dim=128, n_clusters=4096, seed=DB_SEED, mu=None):
    """uint8 [n][dim]: database vectors i0 .. i0+n-1"""
    if mu is None:
        mu = centres(n_clusters, dim, seed)
    ids = np.arange(i0, i0 + n, dtype=np.uint32)
    g = cluster_of(ids, n_clusters, seed)
    x = mu[g] + _noise(seed, ids, dim, 10)
    return np.clip(x, 0, 255).astype(np.uint8)


def query_vectors(nq, n_db, dim=128, n_clusters=4096, seed=DB_SEED, qseed=QUERY_SEED, mu=None):
    """uint8 [nq][dim] queries = perturbed database vectors; also returns the source ids"""
    if mu is None:
        mu = centres(n_clusters, dim, seed)
    j = np.arange(nq, dtype=np.uint32)
    with np.errstate(over="ignore"):
        src = fmix32(np.uint32(qseed) ^ fmix32(j)) % np.uint32(n_db)
    out = np.empty((nq, dim), np.uint8)
    for s in range(0, nq, 65536):
        ids = src[s:s + 65536]
        g = cluster_of(ids, n_clusters, seed)
        x = np.clip(mu[g] + _noise(seed, ids, dim, 10), 0, 255)
        x = x + _noise(qseed, j[s:s + 65536], dim, 11)
        out[s:s + 65536] = np.clip(x, 0, 255).astype(np.uint8)
    return out, src


# ---- codebook training (deterministic Lloyd; numpy or torch tensors on any device) -------

def _kmeans(x, k, iters, rng):
    """x: float32 [n][d] numpy.  Returns [k][d].  Empty clusters are re-seeded."""
    n = x.shape[0]
    if n == 0:
        return np.zeros((k, x.shape[1]), np.float32)
    cent = x[rng.choice(n, size=k, replace=n < k)].astype(np.float32).copy()
    if n < k:
        cent += rng.normal(0, 1e-3, cent.shape).astype(np.float32)
    x2 = (x * x).sum(1)[:, None]
    for _ in range(iters):
        d = x2 - 2.0 * x @ cent.T + (cent * cent).sum(1)[None, :]
        a = d.argmin(1)
        for c in range(k):
            m = a == c
            if m.any():
                cent[c] = x[m].mean(0)
            else:
                cent[c] = x[rng.integers(n)] + rng.normal(0, 1e-2, x.shape[1]).astype(np.float32)
    return cent


def train_tree(train, p, c1, c2, iters=8, seed=1234):
    """Two-level tree in the reference's layouts (createTree, pqt/ProTree.cu:457-510):
    cb1 [c1][dim] (part j uses columns j*vl..), cb2 [p][c1][c2][vl] holding ABSOLUTE
    level-2 centroids trained on the raw segments of the vectors of each L1 cell."""
    train = np.ascontiguousarray(train, np.float32)
    n, dim = train.shape
    vl = dim // p
    rng = np.random.default_rng(seed)
    cb1 = np.zeros((c1, dim), np.float32)
    cb2 = np.zeros((p, c1, c2, vl), np.float32)
    for part in range(p):
        seg = train[:, part * vl:(part + 1) * vl]
        cent = _kmeans(seg, c1, iters, rng)
        cb1[:, part * vl:(part + 1) * vl] = cent
        d = (seg * seg).sum(1)[:, None] - 2.0 * seg @ cent.T + (cent * cent).sum(1)[None, :]
        a = d.argmin(1)
        for c in range(c1):
            cell = seg[a == c]
            if cell.shape[0] == 0:
                cell = cent[c:c + 1]
            cb2[part, c] = _kmeans(cell, c2, iters, rng)
            if cell.shape[0] < c2:  # keep centroids distinct
                cb2[part, c] += rng.normal(0, 0.05, (c2, vl)).astype(np.float32)
    return cb1, cb2
