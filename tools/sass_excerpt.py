#!/usr/bin/env python
"""SASS evidence for profiles/: mnemonic histogram of one kernel plus the instructions around its
TMA / mbarrier use and its hottest loop (the longest backward branch span that contains the
most LDS/LDG).  usage: python tools/sass_excerpt.py <mangled-name substring> [max_lines]"""
import collections
import re
import subprocess
import sys


def main(pat, max_lines=160, lib="product-quantization-tree_b200/libpqt_b200.so"):
    max_lines = int(max_lines)
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    on, name, ins = False, None, []
    for line in out.splitlines():
        s = line.strip()
        if s.startswith("Function :"):
            on = pat in s and name is None
            if on:
                name = s.split(":", 1)[1].strip()
            continue
        if on:
            m = re.match(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", s)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
    if not ins:
        raise SystemExit("no function matching %r" % pat)
    print("# %s" % name)
    print("# %d instructions; cuobjdump -sass %s" % (len(ins), lib))
    hist = collections.Counter()
    for _, t in ins:
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        hist[op.split(".")[0]] += 1
    print("# mnemonics: " + ", ".join("%s %d" % kv for kv in hist.most_common(28)))
    special = [i for i, (_, t) in enumerate(ins) if re.search(r"UBLKCP|SYNCS|UTMA|ELECT|FENCE|MEMBAR", t)]
    if special:
        print("\n# --- TMA bulk copies / mbarrier (UBLKCP, SYNCS) ---")
        shown = set()
        for i in special:
            for j in range(max(0, i - 2), min(len(ins), i + 3)):
                if j not in shown:
                    shown.add(j)
                    print("/*%04x*/  %s ;" % ins[j])
    addr = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                body = ins[addr[tgt]:i + 1]
                w = sum(1 for _, x in body if re.search(r"\bLDS|\bLDG|\bSHFL", x))
                if best is None or w > best[0]:
                    best = (w, addr[tgt], i)
    if best:
        _, lo, hi = best
        print("\n# --- hottest loop: /*%04x*/ .. /*%04x*/, %d instructions (%d shown) ---" %
              (ins[lo][0], ins[hi][0], hi - lo + 1, min(max_lines, hi - lo + 1)))
        body = ins[lo:hi + 1]
        h2 = collections.Counter((t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0] for _, t in body)
        print("# loop mnemonics: " + ", ".join("%s %d" % kv for kv in h2.most_common(16)))
        for a, t in body[:max_lines]:
            print("/*%04x*/  %s ;" % (a, t))


if __name__ == "__main__":
    main(*sys.argv[1:])
