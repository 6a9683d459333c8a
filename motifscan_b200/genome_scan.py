"""Genome-wide scan (BASELINE.json configs[3]): every motif on both strands at every position of
every chromosome, chunked and sharded over the GPUs of a box.

The reference would scan a chromosome as one string (`c_scan_motif` over a list with one sequence
per chromosome, cscore.c:336-389).  Here a chromosome is cut into chunks of `chunk_bp` window
starts, each fetched with a right overlap of (longest motif - 1) bases; the device drops sites
whose start lies in the overlap (`msb_seqs_set_start_limit`), so every window is reported exactly
once, by the chunk that holds its start, and the union equals the unchunked scan bit for bit.
Chunks are dealt to ranks longest-first (`shard.assign_lpt`); ranks never exchange data -- the only
cross-rank step is the caller's gather of the per-motif counts / site arrays.
"""
import ctypes

import numpy as np

from . import _lib, engine, shard

_STRAND_ARG = {"+": 1, "-": 2, "both": 3}


class GenomeSites:
    """Sites of a genome-wide scan in the reference's order: motif, chromosome (sorted names, the
    order of `Genome.chroms`), start, forward before reverse."""

    def __init__(self, chroms, n_motifs, counts, motif, chrom_idx, start, score, strand):
        self.chroms = list(chroms)
        self.n_motifs = n_motifs
        self.counts = counts
        self.motif, self.chrom_idx, self.start, self.score, self.strand = motif, chrom_idx, start, score, strand

    def __len__(self):
        return int(self.counts.sum())


def plan_chunks(chrom_sizes, chunk_bp, halo, world=1, rank=0):
    """This rank's chunks as (chrom, start, end, fetch_end), in (chromosome, start) order."""
    chunks = shard.genome_chunks(chrom_sizes, chunk_bp, halo)
    if world > 1:
        owner, _ = shard.assign_lpt([c[2] - c[1] for c in chunks], world)
        chunks = [c for c, o in zip(chunks, owner) if o == rank]
    return chunks


def scan_genome(genome, pwms, p_value="1e-4", strand="both", chunk_bp=1 << 22, batch_bp=1 << 28,
                ctx=None, world=1, rank=0, collect_sites=True, cutoffs=None):
    """Scan all chromosomes of `genome` (anything with `.chroms`, `.chrom_sizes`, `.fetch_bytes`).

    Returns a GenomeSites for this rank's chunks (`collect_sites=False`: counts only, the site
    arrays stay empty and nothing but the counts leaves the device)."""
    if cutoffs is None:
        cutoffs = [pwm.cutoffs[p_value] for pwm in pwms]
    matrices = [getattr(pwm, "matrix", pwm) for pwm in pwms]
    lmax = max(np.asarray(m).shape[1] for m in matrices)
    ctx = ctx or engine.default_context(0)
    chroms = list(genome.chroms)
    chrom_id = {c: i for i, c in enumerate(chroms)}
    sizes = {c: genome.chrom_sizes[c] for c in chroms}
    chunks = plan_chunks(sizes, chunk_bp, lmax - 1, world, rank)
    motifs = engine.MotifSet(ctx, matrices, cutoffs)
    counts = np.zeros(len(matrices), dtype=np.int64)
    parts = []
    # one pinned staging buffer for the ASCII of a batch: chunks are copied into it once and go to
    # the device at full PCIe speed
    cap = max([f - a for _, a, _, f in chunks] + [1])
    cap = max(cap, min(batch_bp, sum(f - a for _, a, _, f in chunks)))
    pinned = ctypes.c_void_p()
    _lib.check(ctx._lib.msb_pinned_alloc(int(cap), ctypes.byref(pinned)))
    stage = np.ctypeslib.as_array(ctypes.cast(pinned, ctypes.POINTER(ctypes.c_uint8)), shape=(cap,))
    try:
        i = 0
        while i < len(chunks):
            batch, total = [], 0
            while i < len(chunks) and (not batch or total + chunks[i][3] - chunks[i][1] <= batch_bp):
                batch.append(chunks[i])
                total += chunks[i][3] - chunks[i][1]
                i += 1
            off = np.zeros(len(batch) + 1, dtype=np.int64)
            for k, (c, a, _, f) in enumerate(batch):
                piece = np.frombuffer(genome.fetch_bytes(c, a, f), dtype=np.uint8)
                off[k + 1] = off[k] + piece.size
                stage[off[k]:off[k + 1]] = piece
            sset = engine.SequenceSet(ctx, blob=stage[:off[-1]], seq_off=off)
            try:
                sset.set_start_limit([e - a for _, a, e, _ in batch])
                if collect_sites:
                    res = engine.scan(ctx, motifs, sset, _STRAND_ARG[strand])
                    counts += res.counts
                    cid = np.array([chrom_id[c] for c, _, _, _ in batch], dtype=np.int32)
                    base = np.array([a for _, a, _, _ in batch], dtype=np.int64)
                    parts.append((np.repeat(np.arange(len(matrices), dtype=np.int32), res.counts),
                                  cid[res.seq_idx], base[res.seq_idx] + res.start, res.score.copy(), res.strand.copy()))
                    res.close()
                else:
                    engine.scan_device(ctx, motifs, sset, _STRAND_ARG[strand])
                    counts += ctx.site_counts(len(matrices))
            finally:
                sset.close()
    finally:
        motifs.close()
        ctx._lib.msb_pinned_free(pinned)
    if parts:
        motif, cidx, start, score, strnd = (np.concatenate(x) for x in zip(*parts))
        order = np.lexsort((strnd, start, cidx, motif))
        motif, cidx, start, score, strnd = motif[order], cidx[order], start[order], score[order], strnd[order]
    else:
        motif, cidx = np.zeros(0, np.int32), np.zeros(0, np.int32)
        start, score, strnd = np.zeros(0, np.int64), np.zeros(0, np.float64), np.zeros(0, np.int8)
    return GenomeSites(chroms, len(matrices), counts, motif, cidx, start, score, strnd)
