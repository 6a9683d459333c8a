"""Sharding of the scan path over the GPUs of one box.

Every (motif, sequence, window) is independent (cscore.c:336-389 carries no cross-window state),
so work is split by sequence with no exchange step: one process per GPU scans its shard, and the
only cross-rank step is the final host-side gather of per-motif counts and sites.  Regions
(configs[1], [4]) are split in contiguous blocks; a genome (configs[3]) is cut into fixed-size
chunks, each carrying a right halo of `max_motif_len - 1` bases so that a window belongs to the
chunk containing its START, and the chunks are dealt to ranks longest-processing-time-first
(whole chromosomes are too unbalanced: 249 Mbp vs 48 Mbp).
"""
import numpy as np


def region_block(n_items, world, rank):
    """Contiguous block [a, b) of `n_items` for `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_items, world)
    a = rank * base + min(rank, extra)
    return a, a + base + (1 if rank < extra else 0)


def genome_chunks(chrom_sizes, chunk_bp, halo):
    """Cut chromosomes into chunks of `chunk_bp` window starts.  Returns a list of
    (chrom, start, end, fetch_end): windows starting in [start, end) belong to the chunk, which
    needs the bases [start, fetch_end) with fetch_end = min(end + halo, chrom size)."""
    out = []
    for chrom, size in chrom_sizes.items():
        for start in range(0, size, chunk_bp):
            end = min(start + chunk_bp, size)
            out.append((chrom, start, end, min(end + halo, size)))
    return out


def assign_lpt(costs, world):
    """Longest-processing-time-first assignment of items with `costs` to `world` ranks.
    Returns (owner array, per-rank load)."""
    costs = np.asarray(costs, dtype=np.float64)
    owner = np.zeros(len(costs), dtype=np.int64)
    load = np.zeros(world, dtype=np.float64)
    for i in np.argsort(-costs, kind="stable"):
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += costs[i]
    return owner, load


def drop_halo_sites(seq_idx, start, motif_len_of_site, own_len):
    """Mask of sites whose START lies in the chunk's own range [0, own_len[seq]) -- sites that
    start inside the halo belong to the next chunk."""
    return start < own_len[seq_idx]


def gather_sites(local, n_motifs, dist=None, dst=0):
    """Gather per-rank site arrays on rank `dst` and restore the reference's order.

    `local` = dict(counts int64[n_motifs], motif int32[T], seq int64[T] (GLOBAL sequence ids),
    start int64[T], score float64[T], strand int8[T]).  Sequence ids must be global and each
    sequence must live on exactly one rank, so sorting the concatenation by (motif, seq, start,
    strand) reproduces the unsharded list order (cscore.c:336-389).  Returns the merged dict on
    `dst`, None elsewhere.  With dist=None (single process) it only re-sorts."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        parts = [local]
    else:
        parts = [None] * dist.get_world_size() if dist.get_rank() == dst else None
        dist.gather_object(local, parts, dst=dst)
        if dist.get_rank() != dst:
            return None
    motif = np.concatenate([p["motif"] for p in parts])
    seq = np.concatenate([p["seq"] for p in parts])
    start = np.concatenate([p["start"] for p in parts])
    score = np.concatenate([p["score"] for p in parts])
    strand = np.concatenate([p["strand"] for p in parts])
    order = np.lexsort((strand, start, seq, motif))
    counts = np.sum([p["counts"] for p in parts], axis=0)
    return dict(counts=counts, motif=motif[order], seq=seq[order], start=start[order],
                score=score[order], strand=strand[order])


def result_to_local(res, n_motifs, seq_global_ids):
    """engine.ScanResult -> the dict gather_sites expects (sequence ids mapped to global ids)."""
    motif = np.repeat(np.arange(n_motifs, dtype=np.int32), res.counts)
    return dict(counts=res.counts.copy(), motif=motif,
                seq=np.asarray(seq_global_ids, dtype=np.int64)[res.seq_idx],
                start=res.start.astype(np.int64), score=res.score.copy(), strand=res.strand.copy())
