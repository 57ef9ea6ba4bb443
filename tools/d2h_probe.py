#!/usr/bin/env python
"""Device-to-host copy bandwidth per rank, one rank at a time and all ranks together (pinned
host buffers, torch.distributed over NCCL for the barriers).  Explains the end-to-end numbers of
bench.py --gpus N: every rank returns QN/N x k x 8 bytes per step over its own PCIe link.
usage: python -m torch.distributed.run --nproc-per-node N tools/d2h_probe.py [--mb 41]"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=41, help="MB per copy (41 = 1250 queries x 4096 x 8 B)")
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    n = a.mb << 20
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    dst = torch.empty(n, dtype=torch.uint8).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def run():
        t0 = time.perf_counter()
        for _ in range(a.reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        return n * a.reps / (time.perf_counter() - t0) / 1e9

    run()
    alone = []
    for r in range(world):
        barrier()
        bw = run() if r == rank else 0.0
        barrier()
        t = torch.tensor([bw], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        alone.append(float(t[0]))
    barrier()
    bw = run()
    barrier()
    t = torch.zeros(world, dtype=torch.float64, device="cuda")
    t[rank] = bw
    if world > 1:
        dist.all_reduce(t)
    if rank == 0:
        print(json.dumps({"mb_per_copy": a.mb, "ranks": world, "d2h_gbs_alone": [round(x, 1) for x in alone],
                          "d2h_gbs_all_together": [round(float(x), 1) for x in t.tolist()],
                          "numa_note": "pinned by torch (cudaHostAlloc), no explicit NUMA binding"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
