"""Host logic of the multi-GPU path (one process per GPU, torch.distributed).

The bin-ordered code array is cut into contiguous slices with equal vector counts
(pqt_set_shard); rank r owns the queries [r*q_per_rank, (r+1)*q_per_rank) of a batch and the
code rows of its slice.  Per batch every rank dispatches the candidates of its own queries to
the shards that hold them (pqt_shard_dispatch), scans its own inbox and stores every distance
into the owner's array (pqt_shard_scan_p2p), and ranks its own queries (pqt_shard_rank); two
stream-ordered barriers separate the phases.  The helpers below restate the partition rules on
the host (tests, sizing)."""
import torch


def shard_bounds(n, rank, world):
    """[lo, hi) of rank's slice of the bin-ordered list (same arithmetic as pqt_set_shard)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_of(pos, n, world):
    """Shard that holds bin-order position `pos`: the rule of dispatch_kernel
    (r = number of shard lower bounds 1..world-1 that are <= pos)."""
    r = 0
    for i in range(1, world):
        r += 1 if pos >= (n * i) // world else 0
    return r


def query_slice(qn, rank, world):
    assert qn % world == 0, "the query batch must divide evenly over the ranks"
    per = qn // world
    return rank * per, (rank + 1) * per


def dispatch(list_pos, n_list, n, world):
    """Reference of dispatch_kernel on torch tensors: list_pos [Q][max_vec] global positions,
    n_list [Q].  Returns per shard r a list over queries of (local positions, entry numbers),
    entries in list order."""
    out = []
    los = torch.tensor([(n * r) // world for r in range(world + 1)], dtype=torch.int64)
    for r in range(world):
        rows = []
        for q in range(list_pos.shape[0]):
            p = list_pos[q, :int(n_list[q])].to(torch.int64)
            m = (p >= los[r]) & (p < los[r + 1])
            rows.append((p[m] - los[r], torch.nonzero(m).flatten()))
        out.append(rows)
    return out
