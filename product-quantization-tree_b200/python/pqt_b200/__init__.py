"""pqt_b200 -- Python host mirror of the reference's query-path interface over the
C ABI (include/pqt_b200.h).  The product is libpqt_b200.so; this package is the
binding tests/ and bench.py use."""
from .capi import (PAD_IDX, Params, PerturbationProTree, PqtError, Stats, build, lib,  # noqa: F401
                   LIB_PATH, EXPORTS)
from . import formats, synth  # noqa: F401
