#!/usr/bin/env python
"""bench.py -- queries/sec of the PQT query hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
  python bench.py --impl reference --gpus N ...            (CPU arm: the oracle port of the
                                                            reference's algorithm on the
                                                            host cores)

A "step" is one pass of queryKNN over one batch of QN synthetic queries.  Default workload
(BASELINE.json configs[2], the configuration the metric is quoted on): 1B x 128-d synthetic
SIFT-shaped uint8 DB, p=4, c1=c2=32, lineparts=32, 10k-query batch, k=4096, HASH_SIZE=4e8.
`--workload c2` selects configs[1] (1M vectors, lineparts=16); --n / --lineparts override.
The index (128 GB of line codes at 1B) is built on the GPU chunk by chunk, as
test/test1B.cpp:783-871 does with SIFT1B; with --gpus N every rank builds only its own
bin-range shard.  Besides queryKNN the 1-B variant queryBIGKNNRerank2 is measured on the same
index ("big_variant").  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "product-quantization-tree_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tools", "synthdb"))

import numpy as np  # noqa: E402

DB_SEED = 20160627
QUERY_SEED = 424242


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c2", "c3"],
                    help="c3 = BASELINE configs[2]: 1B vectors, lineparts 32 (default); "
                         "c2 = configs[1]: 1M vectors, lineparts 16")
    ap.add_argument("--n", "--dbsize", dest="n", type=int, default=0, help="database vectors (overrides the workload)")
    ap.add_argument("--lineparts", type=int, default=0)
    ap.add_argument("--qn", type=int, default=10000, help="queries per batch")
    ap.add_argument("--k", type=int, default=4096)
    ap.add_argument("--c1", type=int, default=32)
    ap.add_argument("--c2", type=int, default=32)
    ap.add_argument("--p", type=int, default=4)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--hashsize", type=int, default=400000000)
    ap.add_argument("--clusters", type=int, default=0,
                    help="cluster centres of the synthetic data (0 = max(4096, n // 256))")
    ap.add_argument("--train", type=int, default=300000)
    ap.add_argument("--chunk", type=int, default=10000000, help="vectors per build chunk (test/test1B.cpp:623)")
    ap.add_argument("--mode", default="shard", choices=["shard", "replica"],
                    help="N>1: 'shard' = index sharded by bin range (BASELINE configs[3]): every rank "
                         "owns a slice of the batch and a slice of the line codes; candidates are "
                         "dispatched to the shard that holds them and the scan results stored into "
                         "the owner's arrays over NVLink; 'replica' = a full index per GPU")
    ap.add_argument("--variants", default="knn,big", help="comma list of knn, big")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cpu-version", action="store_true",
                    help="skip the cpu_version/ (CPU twin, config C1) leg (~2 min of host time, cached per box)")
    ap.add_argument("--cache-dir", default="/dev/shm/pqt_b200_bench",
                    help="where the index files for the CPU arms live (tmpfs: 136 GB at 1B)")
    ap.add_argument("--rm-files", action="store_true",
                    help="delete the index files at exit (default: left for the next arm on this box)")
    a = ap.parse_args()
    if a.n <= 0:
        a.n = 1000000000 if a.workload == "c3" else 1000000
    if a.lineparts <= 0:
        a.lineparts = 32 if a.workload == "c3" else 16
    if a.clusters <= 0:
        a.clusters = max(4096, a.n // 256)
    return a


def workload_name(a):
    return ("%s x %d-d synthetic SIFT-shaped uint8 DB, p=%d c1=%d c2=%d lineparts=%d, %d-query "
            "batch, k=%d, hash=%d" % (a.n, a.dim, a.p, a.c1, a.c2, a.lineparts, a.qn, a.k,
                                      a.hashsize))


def config_of(a):
    """identical in both arms (the driver compares the dicts)"""
    return {"workload": workload_name(a), "call": "queryKNN", "data_seed": DB_SEED,
            "clusters": a.clusters}


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML; the same fields as the
    nvidia-smi clocks line of /opt/skills/guides/B200_PROFILING.md).  NVML is opened and the
    first sample taken synchronously, so that even a few-ms region reports real samples."""

    def __init__(self, index):
        self.index = index
        self.sm = []
        self.reasons = set()
        self.max_sm = None
        self.stop_flag = threading.Event()
        self.thr = None
        self.err = None
        self.nv = None
        self.h = None
        self.names = {}

    def _sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, nme in self.names.items():
            if r & bit:
                self.reasons.add(nme)

    def _run(self):
        try:
            while not self.stop_flag.is_set():
                self._sample()
                time.sleep(0.002)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except ValueError:
                    pass
            self.h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
        except Exception as e:  # NVML missing: report it, never fake a number
            self.err = repr(e)
            return
        self.thr = threading.Thread(target=self._run, daemon=True)
        self.thr.start()

    def stop(self):
        # one more sample while the last kernels are still in flight / just finished
        if self.nv is not None and self.err is None:
            try:
                self._sample()
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
        self.stop_flag.set()
        if self.thr:
            self.thr.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm,
                    "reasons": ["nvml unavailable: %s" % self.err], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------
# inputs shared by both arms (setup, untimed): centre table, codebooks, queries
def build_inputs(a, device):
    """-> dict(mu uint8 [G][dim] on device, cb1, cb2 (numpy), Q8 uint8 [QN][dim] on device)"""
    import torch
    import synthdb
    from pqt_b200 import synth_torch
    mu = synthdb.centres_u8(a.clusters, a.dim, DB_SEED, device)
    if torch.device(device).type == "cuda":
        ntrain = min(a.train, a.n)
        buf = torch.empty((ntrain, a.dim), dtype=torch.uint8, device=device)
        Xt = synthdb.db_u8(buf, 0, ntrain, mu, DB_SEED)
    else:  # CPU-only box: the torch port of the generator
        ntrain = min(a.train, a.n)
        Xt = synth_torch.db_vectors(0, ntrain, a.dim, a.clusters, mu=mu.cpu().numpy().astype(np.int32),
                                    device=device)
    cb1, cb2 = synth_torch.train_tree(Xt.to(torch.float32), a.p, a.c1, a.c2, iters=10)
    Q8, src = synthdb.queries_u8(a.qn, a.n, mu, DB_SEED, QUERY_SEED)
    return dict(mu=mu, cb1=cb1, cb2=cb2, Q8=Q8, src=src)


def cache_paths(a):
    """Where the index files of this workload live: the first candidate directory that already
    holds a complete set, else the first one with room for it (tmpfs first: the CPU arms
    memory-map the files; a disk directory is the fallback on boxes with a small /dev/shm)."""
    import synthdb
    key = "n%d_d%d_p%d_c%d_%d_h%d_g%d" % (a.n, a.dim, a.p, a.c1, a.c2, a.hashsize, a.clusters)
    need = a.n * (a.lineparts * 4 + 4) + 2 * a.hashsize * 4

    def paths_in(base):
        d = os.path.join(base, key)
        p = synthdb.index_files(os.path.join(d, "synth"), a.dim, a.p, a.c1, a.c2, a.lineparts)
        p["dir"] = d
        p["mu"] = os.path.join(d, "mu.u8")
        return p

    bases = [a.cache_dir, "/tmp/pqt_b200_bench", os.path.join(ROOT, "gpurun_out", "pqt_b200_bench")]
    cands = [paths_in(b) for b in dict.fromkeys(bases)]
    for p in cands:
        if synthdb.files_complete(p, a.n, a.hashsize, a.lineparts) and os.path.exists(p["ppqt"]):
            return p
    for p in cands:
        try:
            os.makedirs(p["dir"], exist_ok=True)
            st = os.statvfs(p["dir"])
            if st.f_bavail * st.f_frsize >= need * 1.02:
                return p
        except OSError:
            pass
    return cands[0]


def ensure_index_files(a, inp, device_index):
    """The reference's index files of the synthetic DB (tool_createdb's outputs), written by
    the native tool_synthdb in a child process.  Returns (paths, seconds spent, built?)."""
    import synthdb
    from pqt_b200 import formats
    paths = cache_paths(a)
    if synthdb.files_complete(paths, a.n, a.hashsize, a.lineparts) and os.path.exists(paths["ppqt"]):
        return paths, 0.0, False
    os.makedirs(paths["dir"], exist_ok=True)
    need = a.n * (a.lineparts * 4 + 4) + 2 * a.hashsize * 4
    st = os.statvfs(paths["dir"])
    if st.f_bavail * st.f_frsize < need * 1.02:
        raise RuntimeError("not enough space in %s for the index files (%.1f GB needed)"
                           % (paths["dir"], need / 1e9))
    formats.write_ppqt(paths["ppqt"], a.dim, a.p, inp["cb1"], inp["cb2"])
    inp["mu"].cpu().numpy().tofile(paths["mu"])
    t0 = time.perf_counter()
    synthdb.run_tool(os.path.join(paths["dir"], "synth"), a.n, a.dim, a.p, a.c1, a.c2, a.lineparts,
                     a.hashsize, a.clusters, DB_SEED, paths["mu"], device=device_index,
                     chunksize=a.chunk)
    return paths, time.perf_counter() - t0, True


def load_host_index(paths, a):
    """memory-maps the index files (tmpfs pages, no second copy)"""
    return dict(prefix=np.memmap(paths["prefix"], np.uint32, "r"),
                counts=np.memmap(paths["count"], np.uint32, "r"),
                db_idx=np.memmap(paths["dbIdx"], np.uint32, "r"),
                lines=np.memmap(paths["lines"], np.uint32, "r", shape=(a.n, a.lineparts)))


def remove_index_files(a):
    import shutil
    shutil.rmtree(cache_paths(a)["dir"], ignore_errors=True)


def cpu_version_baseline(a, device):
    """BASELINE config 1: the reference's CPU twin (cpu_version/, compiled unmodified against the
    Eigen stand-in into oracle/_ref/cpu_version_bench) as treequantizer<float,128,16,8,4,4,32>
    on 1M x 128-d synthetic vectors, query(20000, 500) over 1k queries
    (cpu_version/tools/query.cpp:42,133-138): single thread as the reference runs it, and one
    index replica per host core.  Cached per box (the CPU numbers do not depend on the arm)."""
    import subprocess
    import torch
    import synthdb
    from pqt_b200 import formats
    exe = os.path.join(ROOT, "oracle", "_ref", "cpu_version_bench")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/cpu_version_bench is not built (the reference tree was absent at build time)"}
    d = os.path.join(a.cache_dir, "c1_cpu_version")
    cache = os.path.join(d, "result.json")
    if os.path.exists(cache):
        return json.load(open(cache))
    os.makedirs(d, exist_ok=True)
    n, qn, ncl, dim = 1000000, 1000, 4096, 128
    mu = synthdb.centres_u8(ncl, dim, DB_SEED, device)
    X = synthdb.db_u8(torch.empty((n, dim), dtype=torch.uint8, device=device), 0, n, mu, DB_SEED)
    Q8, _ = synthdb.queries_u8(qn, n, mu, DB_SEED, QUERY_SEED)
    _, gt = synthdb.exact_1nn(Q8, n, mu, DB_SEED)
    formats.write_mem(os.path.join(d, "base.umem"), X.cpu().numpy())
    formats.write_mem(os.path.join(d, "query.umem"), Q8.cpu().numpy())
    formats.write_mem(os.path.join(d, "gt.imem"), gt.cpu().numpy().astype(np.int32)[:, None])
    del X
    threads = os.cpu_count() or 1
    out = {"kind": "cpu_version", "config": "BASELINE configs[0]: 1M x 128-d synthetic, p=4 c1=16 c2=8 "
           "(W=4, LP=32), 1k queries, query(20000, 500)", "host_cores": threads}
    for label, nt in (("single_thread", 1), ("all_cores", threads)):
        r = subprocess.run([exe, os.path.join(d, "base.umem"), os.path.join(d, "query.umem"),
                            os.path.join(d, "gt.imem"), "100000", str(qn), str(nt)],
                           capture_output=True, text=True)
        if r.returncode != 0:
            return {"unavailable": "cpu_version_bench failed: " + (r.stderr or r.stdout)[-300:]}
        out[label] = json.loads(r.stdout.strip().splitlines()[-1])
    out["value"] = out["all_cores"]["queries_per_s"]
    out["unit"] = "queries/s"
    out["cores"] = threads
    out["recall_at_1"] = out["single_thread"]["recall_at_1"]
    for f in ("base.umem", "query.umem", "gt.imem"):
        os.remove(os.path.join(d, f))
    json.dump(out, open(cache, "w"))
    return out


def oracle_handles():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pqt_oracle as po
    po.build()
    return po


def cpu_query(a, po, host_index, cb1, cb2, Q, sample, nthreads, big=False):
    Qs = np.ascontiguousarray(Q[:sample])
    t0 = time.perf_counter()
    if big:
        prm = po.big_params(a.dim, a.p, a.c1, a.c2, a.lineparts, hash_size=a.hashsize)
        d, i, _ = po.query_big_knn_rerank2(prm, cb1, cb2, host_index["prefix"], host_index["counts"],
                                           host_index["db_idx"], host_index["lines"], Qs, a.k,
                                           nthreads=nthreads)
    else:
        prm = po.default_params(a.dim, a.p, a.c1, a.c2, a.lineparts, hash_size=a.hashsize)
        d, i = po.query_knn(prm, cb1, cb2, host_index["prefix"], host_index["counts"],
                            host_index["db_idx"], host_index["lines"], Qs, a.k, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return sample / dt, d, i


# --------------------------------------------------------------------------------------
def run_reference(a, rank, world):
    """CPU arm: the oracle port of the reference's queryKNN on all host threads, reading the
    index files tool_synthdb wrote (the way tool_query reads tool_createdb's).  This process
    never maps libpqt_b200.so: the GPU only produces the input files, in a child process."""
    if rank != 0:
        return
    po = oracle_handles()
    import torch
    threads = os.cpu_count() or 1
    device = "cuda:0" if torch.cuda.is_available() else "cpu"
    inp = build_inputs(a, device)
    Q = inp["Q8"].to(torch.float32).cpu().numpy()
    if device == "cpu":
        # no GPU at all (development container): the oracle's own builder on a small DB
        from pqt_b200 import synth
        mu = inp["mu"].numpy().astype(np.int32)
        X = synth.db_vectors(0, a.n, a.dim, a.clusters, mu=mu).astype(np.float32)
        prm = po.default_params(a.dim, a.p, a.c1, a.c2, a.lineparts, hash_size=a.hashsize)
        host_index = po.build_index(prm, inp["cb1"], inp["cb2"], X, k1_build=min(16, a.c1))
        files_s, built = 0.0, True
    else:
        paths, files_s, built = ensure_index_files(a, inp, 0)
        inp["mu"] = None
        torch.cuda.empty_cache()
        host_index = load_host_index(paths, a)
    sample = a.cpu_sample or min(a.qn, 64 * threads)
    for _ in range(a.warmup):
        cpu_query(a, po, host_index, inp["cb1"], inp["cb2"], Q, min(sample, 16 * threads), threads)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cpu_query(a, po, host_index, inp["cb1"], inp["cb2"], Q, sample, threads)
    dt = time.perf_counter() - t0
    qps = sample * a.steps / dt
    out = {
        "impl": "reference", "metric": "queries/sec", "value": qps, "unit": "queries/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1000.0 * dt / a.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(a),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": "first %d of the batch's %d queries per step, all %d host threads "
                                   "(OpenMP over queries), oracle port of queryKNN over the index "
                                   "files (memory-mapped)" % (sample, a.qn, threads)},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "setup": {"index_files_s": files_s, "index_files_built_here": built},
    }
    print(json.dumps(out), flush=True)
    if a.rm_files:
        remove_index_files(a)


# --------------------------------------------------------------------------------------
def build_index_chunked(a, t, inp, rank, world, device):
    """Both passes of the chunked GPU build.  world > 1: pass 1 is split over the ranks (bins
    all-gathered), pass 2 runs over every chunk on every rank but encodes only the vectors of
    the rank's own bin-range slice."""
    import torch
    import torch.distributed as dist
    import synthdb
    n, chunk = a.n, min(a.chunk, a.n)
    buf = torch.empty((chunk, a.dim), dtype=torch.uint8, device=device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    per = (n + world - 1) // world
    per = (per + 3) // 4 * 4
    bins_all = torch.empty((per * world,), dtype=torch.int32, device=device)
    lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
    for i0 in range(lo, hi, chunk):
        m = min(chunk, hi - i0)
        X = synthdb.db_u8(buf, i0, m, inp["mu"], DB_SEED)
        torch.cuda.synchronize()
        t.assignBins(X, m, bins_all[i0:i0 + m])
    if world > 1:
        dist.all_gather_into_tensor(bins_all, bins_all[rank * per:(rank + 1) * per].clone())
        torch.cuda.synchronize()
    t.setDBFromBins(bins_all[:n], n)
    del bins_all
    torch.cuda.empty_cache()
    t1 = time.perf_counter()
    t.lineDistBegin(n, a.lineparts)
    for i0 in range(0, n, chunk):
        m = min(chunk, n - i0)
        X = synthdb.db_u8(buf, i0, m, inp["mu"], DB_SEED)
        torch.cuda.synchronize()
        t.lineDistChunk(X, i0, m)
    t.lineDistEnd()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    del buf
    torch.cuda.empty_cache()
    return {"bins_s": t1 - t0, "lines_s": t2 - t1}


def ground_truth(a, inp, rank, world, device):
    """exact 1-NN ids of the queries (rank-parallel over id ranges)"""
    import torch
    import torch.distributed as dist
    import synthdb
    per = (a.n + world - 1) // world
    lo, hi = min(a.n, rank * per), min(a.n, (rank + 1) * per)
    sc, arg = synthdb.exact_1nn(inp["Q8"], a.n, inp["mu"], DB_SEED, i_lo=lo, i_hi=hi)
    if world > 1:
        scs = [torch.empty_like(sc) for _ in range(world)]
        args = [torch.empty_like(arg) for _ in range(world)]
        dist.all_gather(scs, sc)
        dist.all_gather(args, arg)
        sc_all = torch.stack(scs)     # [world][QN]; ties -> lowest rank = lowest id
        j = sc_all.argmin(0)
        arg = torch.stack(args).gather(0, j[None, :])[0]
    return arg.cpu().numpy().astype(np.uint32)


def run_b200(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import pqt_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = "cuda:%d" % local_rank
    # one explicit stream for everything: torch ops, NCCL (ordered against the current stream)
    # and the library's launches.  The legacy default stream would not do: its handle is NULL,
    # which pqt_set_stream reads as "the handle's own stream".
    stream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(stream)
    variants = [v for v in a.variants.split(",") if v]
    inp = build_inputs(a, device)
    # The CPU legs (cpu_baseline, parity sample) read the reference's index files; they are
    # written first, by the native tool in a child process, while this process holds no index
    # (at 1B either one needs most of the GPU's memory).
    files = None
    if world == 1 and not a.no_cpu_baseline:
        try:
            files = ensure_index_files(a, inp, local_rank)
        except Exception as e:  # noqa: BLE001
            print("bench.py: index files unavailable (%r); cpu_baseline skipped" % (e,), file=sys.stderr)
    t_setup0 = time.perf_counter()
    if world > 1:  # one set of codebooks for everybody
        cb1 = torch.from_numpy(inp["cb1"]).to(device)
        cb2 = torch.from_numpy(inp["cb2"]).to(device)
        dist.broadcast(cb1, 0)
        dist.broadcast(cb2, 0)
        inp["cb1"], inp["cb2"] = cb1.cpu().numpy(), cb2.cpu().numpy()
    gt = ground_truth(a, inp, rank, world, device)
    torch.cuda.empty_cache()
    gt_s = time.perf_counter() - t_setup0

    sharded = world > 1 and a.mode == "shard"
    replica = world > 1 and a.mode == "replica"
    t = pqt_b200.PerturbationProTree(a.dim, a.p, a.p, local_rank)
    t.set_params(hash_size=a.hashsize, k1_build=min(16, a.c1))
    t.set_stream(stream.cuda_stream)
    t.setTree(inp["cb1"], inp["cb2"])
    if sharded:
        assert a.qn % world == 0, "qn must be divisible by the number of ranks"
        t.setShard(rank, world)
    build = build_index_chunked(a, t, inp, 0 if replica else rank, 1 if replica else world, device)
    QN, k = a.qn, a.k
    Qd = inp["Q8"].to(torch.float32).contiguous()
    Qh = Qd.cpu().pin_memory()
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
        token = torch.zeros(1, dtype=torch.int32, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def shard_step(Qdev, oi, od):
        # 1. Steps A-E1 for the own queries, the LUT of every query, and the dispatch of the own
        #    candidates to the shards that hold them (peer stores over NVLink)
        t.shardDispatch(Qdev, QN, k, q_lo, q_hi)
        dist.all_reduce(token)  # 2. stream-ordered cross-rank barrier: every inbox is complete
        # 3. scan the own inbox; distances go straight into the owners' arrays (peer stores)
        t.shardScanP2P(QN, k)
        dist.all_reduce(token)  # 4. barrier: every distance has arrived
        t.shardRank(nq_out, k, oi, od)  # 5. rank the own queries

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def measure(big):
        """device-resident and end-to-end timing of one variant; returns a dict"""
        call = t.queryBIGKNNRerank2 if big else t.queryKNN

        def step_device():
            if sharded and not big:
                shard_step(Qd, out_i, out_d)
            else:
                call(Qd[q_lo:q_hi], nq_out, k, out_i, out_d)

        def step_e2e():
            if sharded and not big:
                shard_step(Qh.to(device, non_blocking=True), pin_i, pin_d)
            else:
                call(Qh[q_lo:q_hi], nq_out, k, pin_i, pin_d)

        for _ in range(max(a.warmup, 3)):
            step_device()
        barrier()
        # device-resident timing (value): CUDA events on the launching stream, L2 flushed
        # between steps (flush outside the event pairs); nothing synchronises with the host
        # inside a step
        def timed_pass():
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
            return sum(e0.elapsed_time(e1) for e0, e1 in evs)

        sampler = ClockSampler(local_rank)
        sampler.start()
        dev_ms = timed_pass()
        clocks = sampler.stop()
        # the same pass again with the library's per-stage events on (a host synchronisation
        # after every stage): stage times and counters for the roofline, not part of `value`
        t.profile(True)
        t.reset_stats()
        prof_ms = timed_pass()
        st = t.stats()
        t.profile(False)
        # end-to-end through the public call with host buffers
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
        res_i = pin_i.numpy().view(np.uint32).copy()
        res_d = pin_d.numpy().copy()
        if world > 1:
            gi = [torch.empty_like(out_i) for _ in range(world)]
            gd = [torch.empty_like(out_d) for _ in range(world)]
            dist.all_gather(gi, torch.from_numpy(res_i.view(np.int32)).to(device))
            dist.all_gather(gd, torch.from_numpy(res_d).to(device))
            full_i = torch.cat(gi).cpu().numpy().view(np.uint32)
            full_d = torch.cat(gd).cpu().numpy()
        else:
            full_i, full_d = res_i, res_d
        return dict(dev_ms=float(tm[0]), e2e_ms=float(tm[1]), st=st, clocks=clocks, idx=full_i,
                    dist=full_d, prof_ms=prof_ms)

    results = {}
    for v in variants:
        if v == "big" and (world > 1 or a.p != 4):
            continue  # the 1-B variant runs on unsharded handles (p = 4) only
        results[v] = measure(v == "big")
    # second end-to-end operating point: same 4096-candidate scan, k = 100 results returned
    e2e_k100 = None
    if "knn" in results and world == 1 and k > 100 and mv <= 4096:
        t.set_params(max_vec=mv)
        pi = torch.empty((nq_out, 100), dtype=torch.int32).pin_memory()
        pd = torch.empty((nq_out, 100), dtype=torch.float32).pin_memory()
        for _ in range(2):
            t.queryKNN(Qh[q_lo:q_hi], nq_out, 100, pi, pd)
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            t.queryKNN(Qh[q_lo:q_hi], nq_out, 100, pi, pd)
        torch.cuda.synchronize()
        tm = torch.tensor([(time.perf_counter() - t0) * 1000.0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_k100 = {"value": QN * a.steps / (float(tm[0]) * 1e-3), "unit": "queries/s",
                    "ms_per_step": float(tm[0]) / a.steps, "k": 100, "max_vec": mv,
                    "d2h_bytes_per_step": int(QN * 100 * 8)}
        t.set_params(max_vec=0)
    if rank != 0:
        return

    peak, peak_src = measured_peak_hbm()
    def summarise(r):
        st = r["st"]
        scan_ms = st.ms_scan / max(1, st.scan_launches)
        # sharded: rank 0 counts the candidates of its own queries (1/world of the batch), which
        # is also what its shard scans on average (1/world of every query's candidates)
        cand_per_launch = st.candidates / max(1, st.scan_launches)
        # split pipeline: the scan kernel reads the code rows only (ids are implicit in the
        # bin-ordered layout and read by the ranking kernel); fused: codes + the 4-byte id
        split = st.stream_scan_launches > 0
        bytes_per_cand = 4 * a.lineparts + (0 if split else 4)
        achieved = cand_per_launch * bytes_per_cand / (scan_ms * 1e-3) / 1e9
        gtu = gt.astype(np.uint32)
        return {
            "value": QN * a.steps / (r["dev_ms"] * 1e-3), "ms_per_step": r["dev_ms"] / a.steps,
            "e2e": {"value": QN * a.steps / (r["e2e_ms"] * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": int(QN * a.dim * 4),
                    "d2h_bytes_per_step": int(QN * k * 8), "ms_per_step": r["e2e_ms"] / a.steps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None,
                         "kernel": "adc_inbox_kernel (ADC scan of the own shard, one warp per query, results stored to the owners over NVLink)" if sharded else
                         ("adc_stream_kernel (ADC scan; ranking = rank2_kernel, stage 'sort')" if split
                          else "rerank_kernel (ADC scan + ranking fused)"),
                         "peak_source": peak_src, "ms_per_launch": scan_ms,
                         "candidates_per_launch": cand_per_launch,
                         "bytes_per_candidate": bytes_per_cand,
                         "stage_ms_per_step": {"tables": st.ms_tables / a.steps, "bins": st.ms_bins / a.steps,
                                               "scan": st.ms_scan / a.steps, "sort": st.ms_sort / a.steps},
                         "profiled_ms_per_step": r["prof_ms"] / a.steps},
            "recall_at_1": float((r["idx"][:, 0] == gtu).mean()),
            "recall_at_100": float((r["idx"][:, :100] == gtu[:, None]).any(1).mean()),
            "gpu_launches": int(st.kernel_launches),
            "exact_rank_queries_per_step": st.exact_rank_queries / a.steps,
            "tie_resolved_queries_per_step": st.tie_resolved_queries / a.steps,
            "clocks": r["clocks"],
        }

    summ = {v: summarise(r) for v, r in results.items()}
    try:  # DRAM bytes per launch from the committed ncu --set full capture of this workload
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%d_lp%d" % (a.n, a.lineparts)
        if "knn" in summ and not sharded:
            kern = "adc_stream_kernel" if results["knn"]["st"].stream_scan_launches > 0 else "rerank_kernel"
            summ["knn"]["roofline"]["traffic"] = tj[kern].get(key, {}).get("bytes")
    except Exception:  # noqa: BLE001
        pass

    # ---- CPU baseline + parity (oracle port over the index files), rank 0
    cpu = None
    parity = {}
    if not a.no_cpu_baseline:
        po = oracle_handles()
        threads = os.cpu_count() or 1
        import synthdb
        paths = files[0] if files else cache_paths(a)  # N > 1: files left by an earlier arm, if any
        have = synthdb.files_complete(paths, a.n, a.hashsize, a.lineparts)
        if have:
            host_index = load_host_index(paths, a)
            sample = a.cpu_sample or min(QN, 16 * threads)
            Qn = Qh.numpy()
            for v, r in results.items():
                qps, d0, i0 = cpu_query(a, po, host_index, inp["cb1"], inp["cb2"], Qn, sample, threads,
                                        big=(v == "big"))
                parity[v] = bool(np.array_equal(i0, r["idx"][:sample]) and
                                 np.array_equal(d0, r["dist"][:sample]))
                if v == "knn" or cpu is None:
                    cpu = {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                           "sample": "first %d of the %d queries, all %d host threads (OpenMP over "
                                     "queries), oracle port of the reference's %s over the index files"
                                     % (sample, QN, threads, "queryKNN" if v == "knn" else "queryBIGKNNRerank2")}
            del host_index
            if a.rm_files:
                remove_index_files(a)

    cpuv = None
    if world == 1 and not a.no_cpu_baseline and not a.no_cpu_version:
        try:
            cpuv = cpu_version_baseline(a, device)
        except Exception as e:  # noqa: BLE001
            cpuv = {"unavailable": repr(e)}
    head = summ.get("knn") or next(iter(summ.values()))
    par = ("bin-range shards x%d: candidates dispatched to the shard that holds them, scan fused with the "
           "exchange of its results over peer memory (NVLink), two stream-ordered NCCL barriers per batch" % world) if sharded \
        else ("replicas x%d, batch split over the ranks" % world) if replica else "single GPU"
    out = {
        "metric": "queries/sec", "value": head["value"], "unit": "queries/s",
        "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": head["ms_per_step"], "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(a),
        "parallelism": par,
        "l2": "flushed between steps (256 MiB write)",
        "roofline": head["roofline"],
        "cpu_baseline": cpu,
        "cpu_version_baseline": cpuv,
        "e2e": head["e2e"],
        "e2e_k100": e2e_k100,
        "gpu_launches": head["gpu_launches"],
        "clocks": head["clocks"],
        "recall_at_1": head["recall_at_1"],
        "recall_at_100": head["recall_at_100"],
        "exact_rank_queries_per_step": head["exact_rank_queries_per_step"],
        "tie_resolved_queries_per_step": head["tie_resolved_queries_per_step"],
        "parity_vs_oracle_on_cpu_sample": parity.get("knn"),
        "big_variant": None,
        "setup": {"ground_truth_s": gt_s, "index_files_s": files[1] if files else None,
                  "index_build_s": build["bins_s"] + build["lines_s"],
                  "bins_s": build["bins_s"], "lines_s": build["lines_s"],
                  "chunk": min(a.chunk, a.n)},
    }
    if sharded:
        # what bounds the step at this N (per-rank stage times of rank 0, profiled pass)
        st_ms = head["roofline"]["stage_ms_per_step"]
        stage, ms = max(st_ms.items(), key=lambda kv: kv[1])
        per = QN // world
        why = {"tables": "Steps A-C of the own queries plus Step B of every query of the batch",
               "bins": "bin walk + dispatch of the own queries: one CTA per query, %d queries = %.1f waves of "
                       "latency-bound CTAs" % (per, per / (148 * 5.0)),
               "scan": "inbox scan: every shard touches all %d queries (LUT load + pipeline ramp per query for "
                       "about max_vec/%d entries each)" % (QN, world),
               "sort": "ranking of the own queries: one CTA per query, %d queries = %.1f waves" % (per, per / (148 * 4.0))}
        out["limiter"] = {
            "largest_stage": stage, "ms": ms, "why": why.get(stage, ""),
            "outside_the_stages_ms": max(0.0, head["ms_per_step"] - sum(st_ms.values())),
            "outside_the_stages": "two stream-ordered NCCL all-reduce barriers and launch gaps",
            "e2e": "every rank returns %d MB per step over PCIe; with all ranks copying at once the host side "
                   "delivers far less than the 55 GB/s a single rank sees (tools/d2h_probe.py)" % (per * k * 8 >> 20)}
    if "big" in summ:
        b = summ["big"]
        out["big_variant"] = {
            "call": "queryBIGKNNRerank2", "value": b["value"], "unit": "queries/s",
            "ms_per_step": b["ms_per_step"], "e2e": b["e2e"], "roofline": b["roofline"],
            "recall_at_1": b["recall_at_1"], "recall_at_100": b["recall_at_100"],
            "gpu_launches": b["gpu_launches"],
            "parity_vs_oracle_on_cpu_sample": parity.get("big")}
    print(json.dumps(out), flush=True)


def main():
    a = parse()
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
