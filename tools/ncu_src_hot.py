#!/usr/bin/env python
"""Per-CUDA-source-line instruction counts of one kernel in an .ncu-rep (needs -lineinfo).
usage: python tools/ncu_src_hot.py prof.ncu-rep <kernel-regex> [top]"""
import csv
import io
import subprocess
import sys


def main(rep, kern, top=30):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source",
                          "cuda,sass", "--kernel-name", "regex:" + kern], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    cur = None
    lines = []
    hdr = None
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) > 8 and r[0].isdigit() and r[2] == "-":
            ie = hdr.index("Instructions Executed")
            sm = hdr.index("# Samples")
            try:
                lines.append((int(r[ie] or 0), int(r[sm] or 0), cur, int(r[0]), r[1].strip()))
            except ValueError:
                pass
    tot = sum(l[0] for l in lines)
    stot = sum(l[1] for l in lines)
    print("kernel", kern, "executed", tot, "samples", stot)
    for n, s, f, ln, src in sorted(lines, reverse=True)[:top]:
        print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * n / max(1, tot), 100.0 * s / max(1, stot), f, ln, src[:90]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
