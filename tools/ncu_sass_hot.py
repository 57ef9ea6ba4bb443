#!/usr/bin/env python
"""Aggregate the SASS page of one kernel in an .ncu-rep: instructions executed per
opcode and per window of the instruction stream (to locate hot loops).
usage: python tools/ncu_sass_hot.py prof.ncu-rep <kernel-regex> [window]"""
import collections
import csv
import io
import subprocess
import sys


def main(rep, kern, window=300):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name",
                          "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[hi]
    ie, src, smp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    data = []
    for r in rows[hi + 1:]:
        if len(r) <= ie or not r[ie].isdigit():
            continue
        s = r[src].strip()
        toks = s.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        data.append((int(r[ie]), op.split(".")[0], int(r[smp] or 0), s))
    tot = sum(d[0] for d in data)
    stot = sum(d[2] for d in data)
    byop = collections.Counter()
    bys = collections.Counter()
    for n, op, sm, _ in data:
        byop[op] += n
        bys[op] += sm
    print("kernel", kern, "SASS instrs", len(data), "executed", tot, "samples", stot)
    for op, n in byop.most_common(16):
        print("  %-10s %12d %5.1f%%   samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * bys[op] / max(1, stot)))
    print("  window  executed%  samples%  first instruction")
    for i in range(0, len(data), window):
        n = sum(d[0] for d in data[i:i + window])
        s = sum(d[2] for d in data[i:i + window])
        if n * 100.0 / tot > 1.0:
            print("  %6d  %7.1f%%  %7.1f%%  %s" % (i, 100.0 * n / tot, 100.0 * s / max(1, stot), data[i][3][:60]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 300)
