#!/usr/bin/env python3
"""Counters and phase times of the configs[1] block of bench.py (50,000 windows of the resident chr1)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from motifscan_b200 import engine

ctx = engine.Context(0)
pwms = bench.motif_workload()
cut = bench.scan_cutoffs(bench.cutoffs_gpu(engine, ctx, pwms))
pg = bench.make_genome(0, 1.0)
chrom = max(pg.chroms, key=lambda c: pg.chrom_sizes[c])
size = pg.chrom_sizes[chrom]
b0 = int(pg.block_off[pg.chrom_index[chrom]])
resident = engine.SequenceSet.from_packed(ctx, [size], *pg.planes(b0, b0 + (size + 31) // 32))
rng = np.random.default_rng(50)
starts = np.sort(rng.integers(0, size - 1000, size=50000)).astype(np.int64)
w = resident.extract(np.zeros(50000, np.int32), starts, starts + 1000)
motifs = engine.MotifSet(ctx, pwms, cut)
for _ in range(5):
    n = engine.scan_device(ctx, motifs, w, 3)
print(n, ctx.counters(), ctx.timings())
nc = resident.window_ncount(np.zeros(50000, np.int32), starts, 1000)
print("windows with N:", int((nc > 0).sum()), "all N:", int((nc == 1000).sum()), "partial:", int(((nc > 0) & (nc < 1000)).sum()))
