#!/usr/bin/env python3
"""A/B of the prefilter's tile order (ctx option tc_interleave) on configs[1]: phase times, same sites."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from motifscan_b200 import engine, synth

pwms = bench.motif_workload()
blob, off = synth.peak_set(50000, 1000, seed=50)
out = {}
for inter in (0, 1, 0, 1):
    ctx = engine.Context(0)
    ctx.set_option("tc_interleave", inter)
    cut = bench.scan_cutoffs(bench.cutoffs_gpu(engine, ctx, pwms))
    motifs = engine.MotifSet(ctx, pwms, cut)
    sset = engine.SequenceSet(ctx, blob=blob, seq_off=off)
    ts = []
    for _ in range(8):
        n = engine.scan_device(ctx, motifs, sset, 3)
        ts.append(ctx.timings())
    t = {k: round(float(np.median([x[k] for x in ts[3:]])), 4) for k in ("prefilter", "exact", "order")}
    res = engine.scan(ctx, motifs, sset, 3)
    sig = (res.n_sites, int(res.start.astype(np.int64).sum()), float(res.score.sum()))
    print("tc_interleave", inter, t, ctx.counters()["candidates"], sig)
    res.close(), sset.close(), motifs.close(), ctx.close()
