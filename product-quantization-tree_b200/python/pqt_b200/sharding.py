"""Host logic of the multi-GPU path (one process per GPU, torch.distributed).

The index is cut into contiguous slices of the bin-ordered vector list (pqt_set_shard);
every rank scans only its own candidates and marks the rest with (+inf, INT32_MIN).
Since each candidate slot has exactly one owner, an element-wise float MIN / int32 MAX
across ranks assembles exactly the single-GPU candidate arrays; the exchange is one
reduce-scatter per array (NCCL), each rank keeping the queries it will rank.
"""
import torch
import torch.distributed as dist

NOT_MINE_IDX = -(1 << 31)  # 0x80000000 as int32


def shard_bounds(n, rank, world):
    """[lo, hi) of rank's slice of the bin-ordered list (same arithmetic as pqt_set_shard)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_of(pos, n, world):
    """Shard that holds bin-order position `pos`: the rule of the pull-mode scan kernel
    (rerank_kernel<..., PULL>: r = number of shard lower bounds 1..world-1 that are <= pos)."""
    r = 0
    for i in range(1, world):
        r += 1 if pos >= (n * i) // world else 0
    return r


def query_slice(qn, rank, world):
    assert qn % world == 0, "the query batch must divide evenly over the ranks"
    per = qn // world
    return rank * per, (rank + 1) * per


def exchange(val, idx, val_out, idx_out, rank, world, group=None):
    """val/idx: [QN][max_vec] per-shard candidate arrays (float32 / int32).  Fills
    val_out/idx_out [QN/world][max_vec] with the assembled arrays of this rank's queries."""
    if world == 1:
        val_out.copy_(val)
        idx_out.copy_(idx)
        return
    if dist.get_backend(group) == "nccl":
        dist.reduce_scatter_tensor(val_out, val, op=dist.ReduceOp.MIN, group=group)
        dist.reduce_scatter_tensor(idx_out, idx, op=dist.ReduceOp.MAX, group=group)
    else:  # gloo (CPU tests): all-reduce then keep the own slice
        v = val.clone()
        i = idx.clone()
        dist.all_reduce(v, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(i, op=dist.ReduceOp.MAX, group=group)
        lo, hi = query_slice(val.shape[0], rank, world)
        val_out.copy_(v[lo:hi])
        idx_out.copy_(i[lo:hi])


def mask_to_shard(val, idx, cand_pos, n_vec, lo, hi, owns_pad):
    """Reference semantics of adc_scan_kernel's ownership rule on full candidate arrays
    (torch, any device): used by the CPU tests to emulate a shard."""
    qn, mv = val.shape
    a = torch.arange(mv, device=val.device)[None, :]
    real = a < n_vec[:, None].to(torch.int64)
    mine = real & (cand_pos >= lo) & (cand_pos < hi)
    keep = mine | (~real if owns_pad else torch.zeros_like(real))
    v = torch.where(keep, val, torch.full_like(val, float("inf")))
    i = torch.where(keep, idx, torch.full_like(idx, NOT_MINE_IDX))
    return v, i
