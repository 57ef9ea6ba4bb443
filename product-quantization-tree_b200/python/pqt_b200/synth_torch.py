"""torch versions of synth.py (same integer arithmetic, any device) plus k-means
training and exact brute force.  Setup plumbing for tests and bench.py only."""
import numpy as np
import torch

from . import synth

M32 = 0xFFFFFFFF


def fmix32(x):
    x = x & M32
    x = x ^ (x >> 16)
    x = (x * 0x85EBCA6B) & M32
    x = x ^ (x >> 13)
    x = (x * 0xC2B2AE35) & M32
    x = x ^ (x >> 16)
    return x


def _noise(seed, ids, dim, shift):
    base = fmix32((ids + seed) & M32)[:, None]
    d = torch.arange(dim, dtype=torch.int64, device=ids.device)[None, :]
    hh = fmix32((base + d * 0x85EBCA77) & M32)
    s = (hh & 0xFF) + ((hh >> 8) & 0xFF) + ((hh >> 16) & 0xFF) + (hh >> 24)
    return ((s - 510) * 83) >> shift


def cluster_of(ids, n_clusters, seed):
    return fmix32(fmix32(ids) ^ seed) % n_clusters


def db_vectors(i0, n, dim=128, n_clusters=4096, seed=synth.DB_SEED, mu=None, device="cpu",
               chunk=1 << 18):
    """uint8 [n][dim] on `device`, identical to synth.db_vectors"""
    if mu is None:
        mu = synth.centres(n_clusters, dim, seed)
    mu_t = torch.as_tensor(mu, dtype=torch.int64, device=device)
    out = torch.empty((n, dim), dtype=torch.uint8, device=device)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        ids = torch.arange(i0 + s, i0 + e, dtype=torch.int64, device=device)
        g = cluster_of(ids, n_clusters, seed)
        x = mu_t[g] + _noise(seed, ids, dim, 10)
        out[s:e] = x.clamp_(0, 255).to(torch.uint8)
    return out


def query_vectors(nq, n_db, dim=128, n_clusters=4096, seed=synth.DB_SEED,
                  qseed=synth.QUERY_SEED, mu=None, device="cpu"):
    if mu is None:
        mu = synth.centres(n_clusters, dim, seed)
    mu_t = torch.as_tensor(mu, dtype=torch.int64, device=device)
    j = torch.arange(nq, dtype=torch.int64, device=device)
    src = fmix32(fmix32(j) ^ qseed) % n_db
    g = cluster_of(src, n_clusters, seed)
    x = (mu_t[g] + _noise(seed, src, dim, 10)).clamp_(0, 255)
    x = (x + _noise(qseed, j, dim, 11)).clamp_(0, 255)
    return x.to(torch.uint8), src


def _kmeans(x, k, iters, gen):
    n = x.shape[0]
    if n == 0:
        return torch.zeros((k, x.shape[1]), dtype=torch.float32, device=x.device)
    if n >= k:
        sel = torch.randperm(n, generator=gen, device="cpu")[:k].to(x.device)
    else:
        sel = torch.randint(0, n, (k,), generator=gen, device="cpu").to(x.device)
    cent = x[sel].clone()
    if n < k:
        cent += 0.05 * torch.randn(cent.shape, generator=gen).to(x.device)
    for _ in range(iters):
        d = (x * x).sum(1, keepdim=True) - 2.0 * x @ cent.T + (cent * cent).sum(1)[None, :]
        a = d.argmin(1)
        sums = torch.zeros_like(cent).index_add_(0, a, x)
        cnt = torch.zeros(k, device=x.device).index_add_(0, a, torch.ones(n, device=x.device))
        nz = cnt > 0
        cent[nz] = sums[nz] / cnt[nz, None]
        if (~nz).any():
            m = int((~nz).sum())
            r = torch.randint(0, n, (m,), generator=gen, device="cpu").to(x.device)
            cent[~nz] = x[r] + 0.05 * torch.randn((m, x.shape[1]), generator=gen).to(x.device)
    return cent


def train_tree(train, p, c1, c2, iters=10, seed=1234):
    """train: float32 tensor [n][dim] (any device).  Returns numpy cb1 [c1][dim],
    cb2 [p][c1][c2][vl] in the reference's layouts (see synth.train_tree)."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    gen = torch.Generator(device="cpu").manual_seed(seed)
    n, dim = train.shape
    vl = dim // p
    cb1 = torch.zeros((c1, dim), dtype=torch.float32, device=train.device)
    cb2 = torch.zeros((p, c1, c2, vl), dtype=torch.float32, device=train.device)
    for part in range(p):
        seg = train[:, part * vl:(part + 1) * vl].contiguous()
        cent = _kmeans(seg, c1, iters, gen)
        cb1[:, part * vl:(part + 1) * vl] = cent
        d = (seg * seg).sum(1, keepdim=True) - 2.0 * seg @ cent.T + (cent * cent).sum(1)[None, :]
        a = d.argmin(1)
        for c in range(c1):
            cell = seg[a == c]
            if cell.shape[0] == 0:
                cell = cent[c:c + 1]
            cb2[part, c] = _kmeans(cell, c2, iters, gen)
    torch.backends.cuda.matmul.allow_tf32 = prev
    return cb1.cpu().numpy(), cb2.cpu().numpy()


def brute_force_1nn(X_u8, Q_u8, chunk=1 << 17):
    """Exact nearest neighbour of uint8 data: all dot products are integers below 2^24, so
    the fp32 matmul (TF32 off) is exact."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    Q = Q_u8.to(torch.float32)
    qn = (Q * Q).sum(1)
    best = torch.full((Q.shape[0],), float("inf"), device=Q.device)
    arg = torch.zeros(Q.shape[0], dtype=torch.int64, device=Q.device)
    for s in range(0, X_u8.shape[0], chunk):
        X = X_u8[s:s + chunk].to(torch.float32)
        d = qn[:, None] - 2.0 * (Q @ X.T) + (X * X).sum(1)[None, :]
        v, i = d.min(1)
        upd = v < best
        best = torch.where(upd, v, best)
        arg = torch.where(upd, i + s, arg)
    torch.backends.cuda.matmul.allow_tf32 = prev
    return arg
