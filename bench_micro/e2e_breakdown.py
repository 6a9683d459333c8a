#!/usr/bin/env python3
"""Where the end-to-end time of msb_scan_ascii goes (configs[1]): wall time per call against the
device phases, for several upload slice counts."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from motifscan_b200 import _lib, engine, synth

ctx = engine.default_context(0)
_, pwms, _ = synth.motif_set(750, seed=2020)
blob, off = synth.peak_set(50000, 1000, seed=50)
motifs = engine.MotifSet(ctx, pwms)
lmax = max(p.shape[1] for p in pwms)
bblob, boff = synth.background_samples(100000, lmax, seed=1)
bg = engine.SequenceSet(ctx, blob=bblob, seq_off=boff)
motifs.set_cutoffs(np.around(engine.score_select(ctx, motifs, bg, 3, [9])[:, 0], 8))
bg.close()
pin = engine.PinnedArray(blob.size)
pin.array[:] = blob
lib = _lib.load()
for slices in (3, 4, 5, 6):
    ctx.set_option("ascii_slices", slices)
    best = None
    for it in range(6):
        t0 = time.perf_counter()
        res = engine.scan_ascii(ctx, motifs, pin.array, off, 3)
        dt = 1e3 * (time.perf_counter() - t0)
        t = ctx.timings()
        res.close()
        if it >= 2 and (best is None or dt < best[0]):
            best = (dt, t)
    dt, t = best
    dev = t["prefilter"] + t["exact"] + t["order"] + t["d2h"]
    print(f"slices={slices}: wall {dt:.2f} ms | prefilter {t['prefilter']:.2f} exact {t['exact']:.2f} order {t['order']:.2f} "
          f"d2h {t['d2h']:.2f} | unaccounted {dt - dev:.2f}")
t0 = time.perf_counter()
s = engine.SequenceSet(ctx, blob=pin.array, seq_off=off)
t1 = time.perf_counter()
r = engine.scan(ctx, motifs, s, 3)
t2 = time.perf_counter()
print(f"two-step: from_ascii {1e3*(t1-t0):.2f} ms (h2d {ctx.timings()['h2d']:.2f}), scan {1e3*(t2-t1):.2f} ms")
