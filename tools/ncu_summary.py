#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small CSV for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.csv"""
import csv
import io
import subprocess
import sys

WANT = [
    "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
    "smsp__average_warp_latency_issue_stalled_barrier.pct",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.pct",
    "smsp__average_warp_latency_issue_stalled_mio_throttle.pct",
    "smsp__average_warp_latency_issue_stalled_lg_throttle.pct",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(w) for w in WANT if w in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["%s [%s]" % (hdr[c], units[c]) if units[c] else hdr[c] for c in cols])
        for r in rows[2:]:
            w.writerow([r[c] for c in cols])
    print("wrote", out, len(rows) - 2, "kernels")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
