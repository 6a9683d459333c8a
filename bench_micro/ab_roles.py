#!/usr/bin/env python3
"""Phase times of a configs[1] scan with the library named by MSB200_LIB (A/B of compile-time variants)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from motifscan_b200 import engine, synth

pwms = bench.motif_workload()
blob, off = synth.peak_set(50000, 1000, seed=50)
ctx = engine.Context(0)
cut = bench.scan_cutoffs(bench.cutoffs_gpu(engine, ctx, pwms))
motifs = engine.MotifSet(ctx, pwms, cut)
sset = engine.SequenceSet(ctx, blob=blob, seq_off=off)
ts = []
for _ in range(10):
    n = engine.scan_device(ctx, motifs, sset, 3)
    ts.append(ctx.timings())
t = {k: round(float(np.median([x[k] for x in ts[3:]])), 4) for k in ("prefilter", "exact", "order")}
res = engine.scan(ctx, motifs, sset, 3)
print(os.environ.get("MSB200_LIB", "default"), t, (res.n_sites, int(res.start.astype(np.int64).sum())))
