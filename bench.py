#!/usr/bin/env python
"""bench.py -- queries/sec of the PQT query hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
  python bench.py --impl reference --gpus N ...            (CPU arm: the oracle port of the
                                                            reference's algorithm on the
                                                            host cores)

A "step" is one pass of queryKNN over one batch of QN synthetic queries.  Workload
(BASELINE.json configs[1]): 1M x 128-d synthetic SIFT-shaped DB, p=4, c1=c2=32,
lineparts=16, 10k-query batch, k=4096 (the reference's operating point,
test/testPPQT.cpp:285-348), HASH_SIZE=4e8.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "product-quantization-tree_b200", "python"))

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=1000000, help="database vectors")
    ap.add_argument("--qn", type=int, default=10000, help="queries per batch")
    ap.add_argument("--k", type=int, default=4096)
    ap.add_argument("--c1", type=int, default=32)
    ap.add_argument("--c2", type=int, default=32)
    ap.add_argument("--p", type=int, default=4)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--lineparts", type=int, default=16)
    ap.add_argument("--hashsize", type=int, default=400000000)
    ap.add_argument("--clusters", type=int, default=0,
                    help="cluster centres of the synthetic data (0 = max(4096, n // 256))")
    ap.add_argument("--train", type=int, default=300000)
    ap.add_argument("--mode", default="pull", choices=["pull", "shard", "replica"],
                    help="N>1: 'pull' = index sharded by bin range, every rank answers its slice of "
                         "the batch and reads the other shards' line codes over NVLink (no "
                         "collective on the data path); 'shard' = same shards, candidate lists "
                         "all-gathered and the scan results pushed to the query's owner; 'replica' "
                         "= a full index per GPU")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return ("%s x %d-d synthetic SIFT-shaped uint8 DB, p=%d c1=%d c2=%d lineparts=%d, %d-query "
            "batch, k=%d, hash=%d" % (a.n, a.dim, a.p, a.c1, a.c2, a.lineparts, a.qn, a.k,
                                      a.hashsize))


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML, every 5 ms; the
    same fields as the nvidia-smi clocks line of /opt/skills/guides/B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.sm = []
        self.reasons = set()
        self.max_sm = None
        self.stop_flag = threading.Event()
        self.thr = None
        self.err = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            while not self.stop_flag.is_set():
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nme in names.items():
                    if r & bit:
                        self.reasons.add(nme)
                time.sleep(0.005)
        except Exception as e:  # NVML missing: report it, never fake a number
            self.err = repr(e)

    def start(self):
        self.thr = threading.Thread(target=self._run, daemon=True)
        self.thr.start()

    def stop(self):
        self.stop_flag.set()
        if self.thr:
            self.thr.join(timeout=2)
        if self.err or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["nvml unavailable: %s" % self.err],
                    "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------
def build_inputs(a, device):
    """Synthetic DB / queries / codebooks (setup, untimed).  Returns torch tensors."""
    import torch
    from pqt_b200 import synth, synth_torch
    mu = synth.centres(a.clusters, a.dim)
    X8 = synth_torch.db_vectors(0, a.n, a.dim, a.clusters, mu=mu, device=device)
    Q8, src = synth_torch.query_vectors(a.qn, a.n, a.dim, a.clusters, mu=mu, device=device)
    ntrain = min(a.train, a.n)
    cb1, cb2 = synth_torch.train_tree(X8[:ntrain].to(torch.float32), a.p, a.c1, a.c2, iters=10)
    return X8, Q8, src, cb1, cb2


def build_index_gpu(a, X8, cb1, cb2, device_index):
    import torch
    import pqt_b200
    t = pqt_b200.PerturbationProTree(a.dim, a.p, a.p, device_index)
    t.set_params(hash_size=a.hashsize, k1_build=min(16, a.c1))
    t.setTree(cb1, cb2)
    Xf = X8.to(torch.float32).contiguous()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t.buildKBestDB(Xf, a.n)
    t.lineDist(Xf, a.n, a.lineparts)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    del Xf
    return t, build_s


def oracle_handles():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pqt_oracle as po
    po.build()
    return po


def cpu_baseline(a, po, host_index, cb1, cb2, Q, sample, nthreads):
    prm = po.default_params(a.dim, a.p, a.c1, a.c2, a.lineparts, hash_size=a.hashsize)
    Qs = np.ascontiguousarray(Q[:sample])
    t0 = time.perf_counter()
    d, i = po.query_knn(prm, cb1, cb2, host_index["prefix"], host_index["counts"],
                        host_index["db_idx"], host_index["lines"], Qs, a.k, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return sample / dt, d, i


# --------------------------------------------------------------------------------------
def run_reference(a, rank, world):
    """CPU arm: the oracle port of the reference's queryKNN on all host threads."""
    if rank != 0:
        return
    po = oracle_handles()
    import torch
    threads = os.cpu_count() or 1
    use_gpu_setup = torch.cuda.is_available()
    device = "cuda:0" if use_gpu_setup else "cpu"
    X8, Q8, src, cb1, cb2 = build_inputs(a, device)
    Q = Q8.to(torch.float32).cpu().numpy()
    if use_gpu_setup:
        # setup only: the GPU builder produces bit-identical arrays to the oracle's builder
        # (tests/test_gpu_parity.py::test_gpu_builder_matches_oracle_builder)
        t, _ = build_index_gpu(a, X8, cb1, cb2, 0)
        prefix, counts, db_idx = t.getDB()
        lines = t.getLine()
        t.close()
        host_index = dict(prefix=prefix, counts=counts, db_idx=db_idx, lines=lines)
    else:
        prm = po.default_params(a.dim, a.p, a.c1, a.c2, a.lineparts, hash_size=a.hashsize)
        host_index = po.build_index(prm, cb1, cb2, X8.to(torch.float32).numpy(),
                                    k1_build=min(16, a.c1))
    sample = a.cpu_sample or min(a.qn, 512 * threads)
    for _ in range(a.warmup):
        cpu_baseline(a, po, host_index, cb1, cb2, Q, min(sample, 64 * threads), threads)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cpu_baseline(a, po, host_index, cb1, cb2, Q, sample, threads)
    dt = time.perf_counter() - t0
    qps = sample * a.steps / dt
    out = {
        "impl": "reference", "metric": "queries/sec", "value": qps, "unit": "queries/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1000.0 * dt / a.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample_queries_per_step": sample},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": "%d of the batch's %d queries per step, all %d host threads "
                                   "(OpenMP over queries), oracle port of queryKNN" %
                                   (sample, a.qn, threads)},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------
def run_b200(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import pqt_b200
    from pqt_b200 import sharding
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = "cuda:%d" % local_rank
    X8, Q8, src, cb1, cb2 = build_inputs(a, device)
    t, build_s = build_index_gpu(a, X8, cb1, cb2, local_rank)
    stream = torch.cuda.current_stream()
    t.set_stream(stream.cuda_stream)
    QN, k = a.qn, a.k
    Qd = Q8.to(torch.float32).contiguous()
    Qh = Qd.cpu().pin_memory()
    gt = None
    if rank == 0:
        from pqt_b200 import synth_torch
        gt = synth_torch.brute_force_1nn(X8, Q8).cpu().numpy()
    del X8
    sharded = world > 1 and a.mode == "shard"
    pull = world > 1 and a.mode == "pull"
    replica = world > 1 and a.mode == "replica"
    if sharded or pull:
        assert QN % world == 0, "qn must be divisible by the number of ranks"
        t.setShard(rank, world)
    if pull:
        # every rank maps the code slices of the others (CUDA IPC); queries then run through the
        # ordinary public call on each rank's slice of the batch.  If the mapping fails on any
        # rank (peer access unavailable), all ranks fall back to the push pipeline.
        ok = torch.ones(1, dtype=torch.int32, device=device)
        try:
            handles = [None] * world
            dist.all_gather_object(handles, t.shardCodesHandle())
            t.shardCodesOpen(handles)
        except Exception as e:  # noqa: BLE001
            print("rank %d: pull-mode setup failed (%r); falling back to --mode shard" % (rank, e),
                  file=sys.stderr, flush=True)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            pull, sharded = False, True
    mv = t.candidateWidth(k)
    q_lo, q_hi = 0, QN
    if world > 1:
        per = QN // world
        q_lo, q_hi = rank * per, (rank + 1) * per
    nq_out = q_hi - q_lo
    out_i = torch.empty((nq_out, k), dtype=torch.int32, device=device)
    out_d = torch.empty((nq_out, k), dtype=torch.float32, device=device)
    pin_i = torch.empty((nq_out, k), dtype=torch.int32).pin_memory()
    pin_d = torch.empty((nq_out, k), dtype=torch.float32).pin_memory()
    if sharded:
        # fused scan + exchange: candidate arrays of the own queries are owned by the handle
        # and mapped by the peers through CUDA IPC
        t.shardExchangeAlloc(nq_out, mv)
        handles = [None] * world
        dist.all_gather_object(handles, t.shardExchangeHandle())
        t.shardExchangeOpen(handles)
        cand = torch.zeros((QN, mv), dtype=torch.int32, device=device)
        nvec = torch.zeros((QN,), dtype=torch.int32, device=device)
        token = torch.zeros(1, dtype=torch.int32, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def shard_step(Qdev, oi, od):
        # 1. Steps A-E1 for the own queries (+ the LUT of every query)
        t.shardCandidates(Qdev, QN, k, q_lo, q_hi, cand, nvec)
        # 2. everyone learns every query's candidate positions
        dist.all_gather_into_tensor(cand, cand[q_lo:q_hi])
        dist.all_gather_into_tensor(nvec, nvec[q_lo:q_hi])
        # 3. scan the own shard's candidates of all queries, results stored into the owners'
        #    arrays over NVLink; 4. stream-ordered cross-rank barrier; 5. rank the own queries
        t.shardScanP2P(QN, k, cand, nvec)
        dist.all_reduce(token)
        t.shardRank(nvec[q_lo:q_hi], nq_out, k, oi, od)

    def step_device():
        if sharded:
            shard_step(Qd, out_i, out_d)
        else:
            t.queryKNN(Qd[q_lo:q_hi], nq_out, k, out_i, out_d)

    def step_e2e():
        # host buffers in, host buffers out, through the public call
        if sharded:
            shard_step(Qh.to(device, non_blocking=True), pin_i, pin_d)
        else:
            t.queryKNN(Qh[q_lo:q_hi], nq_out, k, pin_i, pin_d)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- warm-up
    for _ in range(max(a.warmup, 3)):
        step_device()
    barrier()
    # ---- device-resident timing (value): CUDA events on the launching stream, L2 flushed
    # between steps (flush outside the event pairs)
    t.profile(True)
    t.reset_stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = []
    barrier()
    for _ in range(a.steps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_device()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    clocks = sampler.stop()
    dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    st = t.stats()
    t.profile(False)
    # ---- end-to-end timing through the public call with host buffers
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    tm = torch.tensor([dev_ms, e2e_s * 1000.0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(tm[0]), float(tm[1])

    res_i = pin_i.numpy().view(np.uint32)
    res_d = pin_d.numpy()
    if world > 1:
        gi = [torch.empty_like(out_i) for _ in range(world)]
        dist.all_gather(gi, torch.from_numpy(res_i.view(np.int32).copy()).to(device))
        full_i = torch.cat(gi).cpu().numpy().view(np.uint32)
    else:
        full_i = res_i
    if rank != 0:
        return

    # ---- roofline of the dominant kernel (ADC scan): algorithmic bytes / event time
    peak, peak_src = measured_peak_hbm()
    bytes_per_cand = 4 * a.lineparts + 4
    scan_ms = st.ms_scan / max(1, st.scan_launches)
    cand_per_launch = st.candidates / max(1, st.scan_launches)
    if sharded:
        cand_per_launch /= world  # each rank scans its slice of the candidates of all queries
    achieved = cand_per_launch * bytes_per_cand / (scan_ms * 1e-3) / 1e9
    traffic = None
    try:  # DRAM bytes per launch from the committed ncu --set full capture of this workload
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if not sharded and a.lineparts == 16:  # the captures are of the lineparts=16 workloads
            traffic = tj["rerank_kernel"].get(str(a.n), {}).get("bytes")
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "kernel": "adc_scan_p2p_kernel" if sharded else "rerank_kernel (ADC scan + ranking fused)",
                "peak_source": peak_src,
                "ms_per_launch": scan_ms, "candidates_per_launch": cand_per_launch,
                "bytes_per_candidate": bytes_per_cand,
                "stage_ms_per_step": {"tables": st.ms_tables / a.steps, "bins": st.ms_bins / a.steps,
                                      "scan": st.ms_scan / a.steps, "sort": st.ms_sort / a.steps}}
    recall1 = float((full_i[:, 0] == gt.astype(np.uint32)).mean())

    # ---- CPU baseline (oracle port) on a bounded sample, rank 0, N = 1 only
    cpu = None
    parity = None
    if world == 1 and not a.no_cpu_baseline:
        po = oracle_handles()
        threads = os.cpu_count() or 1
        prefix, counts, db_idx = t.getDB()
        lines = t.getLine()
        host_index = dict(prefix=prefix, counts=counts, db_idx=db_idx, lines=lines)
        sample = a.cpu_sample or min(QN, 256 * threads)
        qps, d0, i0 = cpu_baseline(a, po, host_index, cb1, cb2, Qh.numpy(), sample, threads)
        parity = bool(np.array_equal(i0, full_i[:sample]) and np.array_equal(d0, res_d[:sample]))
        cpu = {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
               "sample": "first %d of the %d queries, all %d host threads (OpenMP over queries), "
                         "oracle port of the reference's queryKNN" % (sample, QN, threads)}
    out = {
        "metric": "queries/sec", "value": QN * a.steps / (dev_ms * 1e-3), "unit": "queries/s",
        "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": dev_ms / a.steps, "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "l2": "flushed between steps (256 MiB write)",
                   "parallelism": ("bin-range shards x%d, scan fused with peer-memory exchange (NVLink), NCCL all-gather of candidate lists" % world) if sharded
                   else ("bin-range shards x%d, batch split over the ranks, line codes of the other shards read over NVLink inside the fused scan kernel, no collective" % world) if pull
                   else ("replicas x%d, batch split over the ranks" % world) if replica else "single GPU",
                   "index_build_s": build_s},
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": QN * a.steps / (e2e_ms * 1e-3), "unit": "queries/s",
                "h2d_bytes_per_step": int(QN * a.dim * 4 if sharded else nq_out * a.dim * 4) * (1 if sharded else world),
                "d2h_bytes_per_step": int(QN * k * 8), "ms_per_step": e2e_ms / a.steps},
        "gpu_launches": int(st.kernel_launches),
        "clocks": clocks,
        "recall_at_1": recall1,
        "exact_rank_queries_per_step": st.exact_rank_queries / a.steps,
        "tie_resolved_queries_per_step": st.tie_resolved_queries / a.steps,
        "parity_vs_oracle_on_cpu_sample": parity,
    }
    print(json.dumps(out), flush=True)


def main():
    a = parse()
    if a.clusters <= 0:
        a.clusters = max(4096, a.n // 256)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local_rank))
    try:
        run_b200(a, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
