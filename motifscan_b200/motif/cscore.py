"""Drop-in for the reference's extension module `motifscan.motif.cscore`
(reference motifscan/motif/cscore.c:479-494): same two callables, same argument meaning, same
return shapes, computed by libmsb200.so on a B200.

    c_scan_motif(pwms, cutoffs, seqs, strand, n_threads)   -- cscore.c:399-476
    c_score(pwms, seqs, strand, n_threads)                 -- cscore.c:231-302

`n_threads` is accepted and ignored (the GPU kernel has no thread count).  The array-native
variants (`scan_arrays`, `score_array`) skip the per-hit Python objects and are what Scanner uses.
There is no CPU fallback: without the built library or a CUDA device these raise.
"""
import os

from .. import engine

__all__ = ["c_score", "c_scan_motif", "scan_arrays", "score_array"]


def _device():
    return int(os.environ.get("MOTIFSCAN_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def scan_arrays(pwms, cutoffs, seqs, strand, device=None):
    """Scan and return an engine.ScanResult (CSR over motifs; arrays seq_idx, start, score,
    strand in the reference's list order)."""
    ctx = engine.default_context(_device() if device is None else device)
    motifs = engine.MotifSet(ctx, pwms, cutoffs)
    try:
        sset = engine.SequenceSet(ctx, seqs)
        try:
            return engine.scan(ctx, motifs, sset, strand)
        finally:
            sset.close()
    finally:
        motifs.close()


def c_scan_motif(pwms, cutoffs, seqs, strand, n_threads=1):
    """list[n_pwms] of list of [seq_idx:int, start:int, score:float, strand:int(1|2)], ordered by
    (sequence, start, forward before reverse) -- cscore.c:443-471."""
    if strand not in (1, 2, 3):
        # the reference computes nothing for other bit patterns' missing bits; 0 yields no sites
        if strand == 0:
            return [[] for _ in pwms]
        raise ValueError("strand must be 1 (forward), 2 (reverse) or 3 (both)")
    res = scan_arrays(pwms, cutoffs, seqs, strand)
    try:
        seq_idx = res.seq_idx.tolist()
        start = res.start.tolist()
        score = res.score.tolist()
        strd = res.strand.tolist()
        off = res.offsets.tolist()
    finally:
        res.close()
    out = []
    for m in range(len(pwms)):
        a, b = off[m], off[m + 1]
        out.append([list(t) for t in zip(seq_idx[a:b], start[a:b], score[a:b], strd[a:b])])
    return out


def score_array(pwms, seqs, strand, device=None):
    """float64 array (n_pwms, n_seqs): offset-0 window score of every sequence."""
    ctx = engine.default_context(_device() if device is None else device)
    motifs = engine.MotifSet(ctx, pwms, None)
    try:
        sset = engine.SequenceSet(ctx, seqs)
        try:
            return engine.score(ctx, motifs, sset, strand)
        finally:
            sset.close()
    finally:
        motifs.close()


def c_score(pwms, seqs, strand, n_threads=1):
    """list[n_pwms] of list[n_seqs] of float -- cscore.c:282-297.  A sequence shorter than the
    longest motif is undefined behaviour in the reference (cscore.c:195) and a ValueError here."""
    if strand not in (1, 2, 3):
        raise ValueError("strand must be 1 (forward), 2 (reverse) or 3 (both)")
    return score_array(pwms, seqs, strand).tolist()
