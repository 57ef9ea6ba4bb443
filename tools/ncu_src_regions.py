#!/usr/bin/env python
"""Instruction / sample share per source region of one kernel in an .ncu-rep.
usage: python tools/ncu_src_regions.py prof.ncu-rep <kernel-regex> file:lo-hi=name ..."""
import csv
import io
import subprocess
import sys


def main(rep, kern, *regions):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source",
                          "cuda,sass", "--kernel-name", "regex:" + kern], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    cur, hdr, lines = None, None, []
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) > 8 and r[0].isdigit() and r[2] == "-":
            ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
            try:
                lines.append((cur, int(r[0]), int(r[ie] or 0), int(r[sm] or 0)))
            except ValueError:
                pass
    tot = sum(l[2] for l in lines) or 1
    stot = sum(l[3] for l in lines) or 1
    regs = []
    for spec in regions:
        loc, name = spec.split("=")
        f, rng = loc.split(":")
        lo, hi = map(int, rng.split("-"))
        regs.append((f, lo, hi, name))
    acc = {}
    for f, ln, n, s in lines:
        name = "other:" + f
        for rf, lo, hi, rn in regs:
            if f == rf and lo <= ln <= hi:
                name = rn
                break
        a = acc.setdefault(name, [0, 0])
        a[0] += n
        a[1] += s
    print("total inst", tot, "samples", stot)
    for name, (n, s) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        print("%-34s %5.1f%% inst  %5.1f%% samples" % (name, 100.0 * n / tot, 100.0 * s / stot))


if __name__ == "__main__":
    main(*sys.argv[1:])
