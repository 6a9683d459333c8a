"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the MotifScan hot path (reference: motifscan/motif/cscore.c,
motifscan/scanner.py, motifscan/motif/__init__.py:378-401, motifscan/motif/matrix.py,
motifscan/cli/motif.py:119-153).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this package; nothing
under ``motifscan_b200/`` does.

Parity status: PINNED -- see the header of ``oracle/oracle.c`` and ``tests/test_oracle.py``.

Two layers:

* ``liboracle.so`` (``oracle.c``): the native arithmetic (encode, max_raw_score, c_score,
  c_scan_motif semantics) over flat arrays, loaded with ctypes.
* pure-Python restatements of the host-side pieces of the path (regrouping, de-duplication,
  cutoff order statistics, PFM->PWM), written as plain loops.

``load_reference_cscore()`` loads the UNMODIFIED reference extension compiled by
``oracle/Makefile`` into ``oracle/_ref`` (when present).
"""
import ctypes
import glob
import importlib.util
import os
import subprocess
from collections import namedtuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MotifSite = namedtuple("MotifSite", ["start", "score", "strand"])  # scanner.py:16


def build(quiet=True):
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    out = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        c_i32p = ctypes.POINTER(ctypes.c_int32)
        c_i64p = ctypes.POINTER(ctypes.c_int64)
        c_f64p = ctypes.POINTER(ctypes.c_double)
        c_i8p = ctypes.POINTER(ctypes.c_int8)
        lib.orc_max_raw_score.restype = ctypes.c_double
        lib.orc_max_raw_score.argtypes = [c_f64p, ctypes.c_int32]
        lib.orc_encode.restype = None
        lib.orc_encode.argtypes = [ctypes.c_char_p, ctypes.c_int64, c_i8p]
        lib.orc_score.restype = ctypes.c_int
        lib.orc_score.argtypes = [ctypes.c_int32, c_i32p, c_f64p, c_i64p, ctypes.c_int64,
                                  ctypes.c_char_p, c_i64p, ctypes.c_int, c_f64p]
        lib.orc_scan.restype = ctypes.c_int
        lib.orc_scan.argtypes = [ctypes.c_int32, c_i32p, c_f64p, c_i64p, c_f64p, ctypes.c_int64,
                                 ctypes.c_char_p, c_i64p, ctypes.c_int, ctypes.c_int, c_i64p,
                                 ctypes.POINTER(c_i64p), ctypes.POINTER(c_i64p),
                                 ctypes.POINTER(c_f64p), ctypes.POINTER(c_i8p)]
        lib.orc_free.restype = None
        lib.orc_free.argtypes = [ctypes.c_void_p]
        _LIB = lib
    return _LIB


def _ptr(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def flatten_pwms(pwms):
    """list of 4 x L_m matrices -> (lens int32[n], mats float64[sum 4 L], mat_off int64[n])."""
    lens = np.array([len(p[0]) for p in pwms], dtype=np.int32)
    mat_off = np.zeros(len(pwms), dtype=np.int64)
    chunks = []
    at = 0
    for m, p in enumerate(pwms):
        a = np.ascontiguousarray(np.asarray(p, dtype=np.float64))
        assert a.shape == (4, lens[m])
        mat_off[m] = at
        at += a.size
        chunks.append(a.ravel())
    mats = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.float64)
    return lens, mats, mat_off


def flatten_seqs(seqs):
    """list of str/bytes -> (bytes blob, seq_off int64[n+1])."""
    bs = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs])
    return b"".join(bs), off


def encode(seq):
    """cscore.c:81-114."""
    b = seq.encode("utf-8") if isinstance(seq, str) else bytes(seq)
    out = np.empty(len(b), dtype=np.int8)
    _lib().orc_encode(b, len(b), _ptr(out, ctypes.c_int8))
    return out


def max_raw_score(matrix):
    """cscore.c:36-48."""
    a = np.ascontiguousarray(np.asarray(matrix, dtype=np.float64))
    return float(_lib().orc_max_raw_score(_ptr(a, ctypes.c_double), a.shape[1]))


def score_arrays(pwms, seqs, strand):
    """c_score semantics (cscore.c:174-302) -> float64 array (n_pwms, n_seqs)."""
    lens, mats, mat_off = flatten_pwms(pwms)
    blob, seq_off = flatten_seqs(seqs)
    out = np.empty((len(pwms), len(seqs)), dtype=np.float64)
    rc = _lib().orc_score(len(pwms), _ptr(lens, ctypes.c_int32), _ptr(mats, ctypes.c_double),
                          _ptr(mat_off, ctypes.c_int64), len(seqs), blob,
                          _ptr(seq_off, ctypes.c_int64), int(strand), _ptr(out, ctypes.c_double))
    if rc == -1:
        raise ValueError("c_score: a sequence is shorter than the longest motif")
    if rc != 0:
        raise MemoryError()
    return out


def c_score(pwms, seqs, strand, n_threads=1):
    """Same return shape as the reference's c_score: list[n_pwms] of list[n_seqs] of float."""
    return score_arrays(pwms, seqs, strand).tolist()


def scan_arrays(pwms, cutoffs, seqs, strand, n_threads=1):
    """c_scan_motif semantics (cscore.c:317-476) as arrays.

    Returns (counts int64[n_pwms], seq_idx int64[T], start int64[T], score float64[T],
    strand int8[T]) with sites motif-major in the reference's list order.
    """
    lens, mats, mat_off = flatten_pwms(pwms)
    blob, seq_off = flatten_seqs(seqs)
    cut = np.asarray(cutoffs, dtype=np.float64)
    counts = np.zeros(len(pwms), dtype=np.int64)
    p_seq = ctypes.POINTER(ctypes.c_int64)()
    p_start = ctypes.POINTER(ctypes.c_int64)()
    p_score = ctypes.POINTER(ctypes.c_double)()
    p_strand = ctypes.POINTER(ctypes.c_int8)()
    lib = _lib()
    rc = lib.orc_scan(len(pwms), _ptr(lens, ctypes.c_int32), _ptr(mats, ctypes.c_double),
                      _ptr(mat_off, ctypes.c_int64), _ptr(cut, ctypes.c_double), len(seqs), blob,
                      _ptr(seq_off, ctypes.c_int64), int(strand), int(n_threads),
                      _ptr(counts, ctypes.c_int64), ctypes.byref(p_seq), ctypes.byref(p_start),
                      ctypes.byref(p_score), ctypes.byref(p_strand))
    if rc != 0:
        raise MemoryError()
    total = int(counts.sum())
    try:
        seq_idx = np.ctypeslib.as_array(p_seq, shape=(max(total, 1),))[:total].copy()
        start = np.ctypeslib.as_array(p_start, shape=(max(total, 1),))[:total].copy()
        score = np.ctypeslib.as_array(p_score, shape=(max(total, 1),))[:total].copy()
        strd = np.ctypeslib.as_array(p_strand, shape=(max(total, 1),))[:total].copy()
    finally:
        for p in (p_seq, p_start, p_score, p_strand):
            lib.orc_free(ctypes.cast(p, ctypes.c_void_p))
    return counts, seq_idx, start, score, strd


def c_scan_motif(pwms, cutoffs, seqs, strand, n_threads=1):
    """Same return shape as the reference's c_scan_motif (cscore.c:443-471)."""
    counts, seq_idx, start, score, strd = scan_arrays(pwms, cutoffs, seqs, strand, n_threads)
    out = []
    at = 0
    for m in range(len(pwms)):
        n = int(counts[m])
        out.append([[int(seq_idx[k]), int(start[k]), float(score[k]), int(strd[k])]
                    for k in range(at, at + n)])
        at += n
    return out


# ----------------------------------------------------------------------------------------------
# The unmodified reference extension (oracle/_ref), when it has been built.
# ----------------------------------------------------------------------------------------------
def reference_cscore_path():
    hits = sorted(glob.glob(os.path.join(_HERE, "_ref", "cscore*.so")))
    return hits[0] if hits else None


def load_reference_cscore():
    """Load oracle/_ref/cscore*.so directly (avoids importing the reference package, whose
    __init__ chain needs pysam).  Returns the module or None when it was never built."""
    path = reference_cscore_path()
    if path is None:
        return None
    spec = importlib.util.spec_from_file_location("cscore", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ----------------------------------------------------------------------------------------------
# Pure-Python restatements of the host-side steps of the path.
# ----------------------------------------------------------------------------------------------
def make_motif_sites(sites, seq_starts):
    """scanner.py:135-153: pooled per-motif site lists -> [pwm][seq] -> [MotifSite], with
    start shifted by the sequence's genomic start and strand 1/2 mapped to '+'/'-'."""
    out = []
    for sites_pwm in sites:
        per_seq = [[] for _ in seq_starts]
        for seq_idx, pos, score, strand in sites_pwm:
            per_seq[seq_idx].append(
                MotifSite(seq_starts[seq_idx] + pos, score, '+' if strand == 1 else '-'))
        out.append(per_seq)
    return out


def _dedup_one_strand(sites, length):
    """scanner.py:156-168: greedy left-to-right removal of overlapping neighbours."""
    idx = 0
    while idx + 1 < len(sites):
        cur, nxt = sites[idx], sites[idx + 1]
        if nxt.start - cur.start < length:
            if cur.score >= nxt.score:
                del sites[idx + 1]
            else:
                del sites[idx]
        else:
            idx += 1


def deduplicate_motif_sites(motif_sites, lengths):
    """scanner.py:171-193: per (motif, region), per strand separately, then fwd + rev
    concatenated and stably sorted by start."""
    out = []
    for sites_pwm, length in zip(motif_sites, lengths):
        per_seq = []
        for sites in sites_pwm:
            fwd = [s for s in sites if s.strand == '+']
            rev = [s for s in sites if s.strand != '+']
            _dedup_one_strand(fwd, length)
            _dedup_one_strand(rev, length)
            both = fwd + rev
            both.sort(key=lambda s: s.start)
            per_seq.append(both)
        out.append(per_seq)
    return out


def get_score_cutoffs(sampling_scores):
    """motif/__init__.py:378-401.  Index arithmetic is evaluated with Python floats exactly as
    the reference does (`int(n * 0.1 ** e) - 1`)."""
    out = []
    for scores in sampling_scores:
        if len(scores) < 100:
            raise ValueError("each motif must have at least 100 sampling scores")
        n = len(scores)
        n_bits = min(len(str(n)), 7)
        ordered = sorted(scores, reverse=True)
        cut = {}
        for e in range(2, n_bits):
            cut[f"1e-{e}"] = ordered[int(n * 0.1 ** e) - 1]
        out.append(cut)
    return out


def average_cutoffs(cutoffs_all):
    """cli/motif.py:144-153: mean over repeats, np.around(., 8); returns list of dicts."""
    n = len(cutoffs_all[0])
    out = []
    for i in range(n):
        acc = {}
        for rep in cutoffs_all:
            for p, c in rep[i].items():
                acc.setdefault(p, []).append(c)
        out.append({p: np.around(np.mean(v), 8) for p, v in acc.items()})
    return out


def pfm_to_pwm(pfm, bg_freq=None, pseudo=0.001):
    """matrix.py:74-98 (to_ppm), :125-146 (normalize), :148-171 (to_pwm)."""
    pfm = np.asarray(pfm)
    ppm = pfm / pfm.sum(axis=0)
    pseudo_count = pseudo / (1 - 4 * pseudo)
    zero_cols = np.any(ppm == 0, axis=0)
    ppm[:, zero_cols] += pseudo_count
    ppm = ppm / ppm.sum(axis=0)
    if bg_freq is None:
        bg_freq = {b: 0.25 for b in "ACGT"}
    bg = np.asarray([bg_freq[b] for b in "ACGT"]).reshape(4, 1)
    return np.around(np.log(ppm / bg), 5)


def pwm_score(matrix, sequence):
    """matrix.py:216-240 PositionWeightMatrix.score: forward strand, one window, N skipped,
    divided by numpy's max(axis=0).sum() (NOT floored at 0, unlike cscore.c:36-48)."""
    matrix = np.asarray(matrix, dtype=float)
    if len(sequence) != matrix.shape[1]:
        raise ValueError("sequence should have the same length as the PWM")
    rows = {'A': 0, 'C': 1, 'G': 2, 'T': 3}
    raw = 0
    for col, nt in enumerate(sequence.upper()):
        if nt in rows:
            raw += matrix[rows[nt], col]
    return raw / matrix.max(axis=0).sum()
