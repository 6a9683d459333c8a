"""motifscan_b200 -- B200 (sm_100a) implementation of MotifScan's motif-scanning hot path.

Layout
    csrc/            CUDA kernels + the C ABI (include/msb200.h) -> libmsb200.so
    _lib.py          ctypes binding of the C ABI
    engine.py        handle classes (Context, MotifSet, SequenceSet, ScanResult)
    motif/cscore.py  drop-in for the reference's extension module motifscan.motif.cscore
"""
__version__ = "0.1.0"
