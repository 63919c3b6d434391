"""fqcomp28_b200 -- B200-native (sm_100a) implementation of the fqcomp28 codec
hot path behind the C ABI of include/fq28.h.

The product is fqcomp28_b200/libfq28.so (hand-written CUDA kernels) plus the
C++ facade in fqcomp28_b200/host/.  This Python package only binds the C ABI
for bench.py and the parity tests; it has no CPU fallback and never
uses the CPU checker that lives beside the tests.
"""
from .capi import (  # noqa: F401
    ChunkInfo,
    DecArenas,
    EncArenas,
    EncSummary,
    Fq28Error,
    Handle,
    LIB_PATH,
    SYMBOLS,
    load,
)
