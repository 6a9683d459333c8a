"""Score-cutoff computation of `motifscan motif --build` (reference cli/motif.py:101-155) with the
scoring and the order statistics on the device.

Reference arithmetic per repeat i: seed+i -> `genome.random_sequences(n_random, max_length, max_n)`
-> `c_score(matrices, seqs, 3)` -> `get_score_cutoffs` (sort descending, read indices
int(n * 0.1**e) - 1); then the mean over repeats, `np.around(., 8)`.  Here the n_pwms x n_random
score matrix stays on the GPU (`msb_score_select`); only the selected order statistics come back.
The background sampler's legacy-numpy RNG sequence is part of the result (genome/__init__.py:159-176)
and stays on the host; with a `genome.DeviceGenome` only the RNG runs there -- the N count of every
attempt and the sampled windows come out of the resident genome on the device.
"""
import numpy as np

from . import engine
from .motif import cutoff_ranks


def sampled_cutoffs(ctx, motif_set, seqs=None, blob=None, seq_off=None, sset=None):
    """One repeat: {p: cutoff} per motif from one batch of background sequences."""
    sset = sset or engine.SequenceSet(ctx, seqs=seqs, blob=blob, seq_off=seq_off)
    try:
        ranks = cutoff_ranks(sset.n)
        keys = list(ranks)
        sel = engine.score_select(ctx, motif_set, sset, 3, [ranks[k] for k in keys])
    finally:
        sset.close()
    return [{k: float(sel[m, j]) for j, k in enumerate(keys)} for m in range(motif_set.n)]


def build_cutoffs(pwms, genome, n_random=1000000, n_repeat=1, max_n=0, seed=None, ctx=None):
    """Set `pwm.cutoffs` for every PWM like `build_motif` does (cli/motif.py:119-153)."""
    resident = hasattr(genome, "random_sequence_set")     # a genome.DeviceGenome: samples never leave the device
    ctx = genome.ctx if resident else (ctx or engine.default_context(0))
    max_length = max(pwm.length for pwm in pwms)
    motif_set = engine.MotifSet(ctx, [pwm.matrix for pwm in pwms])
    try:
        per_repeat = []
        for i in range(n_repeat):
            s = seed + i if seed is not None else None
            if resident:
                per_repeat.append(sampled_cutoffs(ctx, motif_set, sset=genome.random_sequence_set(n_random, max_length, max_n, s)))
            else:
                seqs = list(genome.random_sequences(n_random, max_length, max_n, s))
                per_repeat.append(sampled_cutoffs(ctx, motif_set, seqs=seqs))
    finally:
        motif_set.close()
    for m, pwm in enumerate(pwms):
        for p in per_repeat[0][m]:
            cutoff = np.around(np.mean([rep[m][p] for rep in per_repeat]), 8)
            pwm.set_cutoff(p_value=p, cutoff=cutoff)
    return pwms
