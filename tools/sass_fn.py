#!/usr/bin/env python
"""Print the SASS of one kernel of the built library.
usage: python tools/sass_fn.py <substring of the mangled name> [lib.so]"""
import subprocess
import sys


def main(pat, lib="product-quantization-tree_b200/libpqt_b200.so"):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    on = False
    for line in out.splitlines():
        if line.strip().startswith("Function :"):
            on = pat in line
        if on:
            print(line)


if __name__ == "__main__":
    main(*sys.argv[1:])
