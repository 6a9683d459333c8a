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


def gather_counts(counts, dist=None, dst=None):
    """The final gather of a sharded scan: per-motif site counts summed over the ranks on the process group's
    host backend (gloo; NCCL is not on this path).  `dst=None`: every rank gets the sum (one all-reduce of
    n_motifs int64).  `dst=r`: the ranks send their counts straight to rank r (one hop each, the lowest-latency
    form over loopback sockets), which returns the sum; the others return None."""
    counts = np.ascontiguousarray(np.asarray(counts, dtype=np.int64))
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return counts.copy()
    import torch
    t = torch.from_numpy(counts.copy())
    if dst is None:
        dist.all_reduce(t)
        return t.numpy()
    if dist.get_rank() == dst:
        parts = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
        dist.gather(t, parts, dst=dst)
        return torch.stack(parts).sum(dim=0).numpy()
    dist.gather(t, None, dst=dst)
    return None


def gather_sites(local, n_motifs, dist=None, dst=0):
    """Gather per-rank site arrays on rank `dst` in the reference's order (cscore.c:336-389: motif, sequence,
    start, forward before reverse).

    `local` = dict(counts int64[n_motifs], seq int64[T] (GLOBAL sequence ids), start int64[T],
    score float64[T], strand int8[T]), motif-major, each rank holding an ASCENDING block of the sequences
    (`region_block`), so motif m's gathered list is rank 0's slice, then rank 1's, ...: the arrays travel as
    plain tensors (`dist.gather`, padded to the largest rank) and are interleaved by
    `engine.merge_motif_major` -- no pickling, no sort.  Returns the merged dict (with `motif`) on `dst`,
    None elsewhere; with dist=None (single process) it returns the input's arrays."""
    from . import engine
    names = ("seq", "start", "score", "strand")
    counts = np.ascontiguousarray(np.asarray(local["counts"], dtype=np.int64))
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        parts_counts = counts[None, :]
        parts = {k: [np.ascontiguousarray(local[k])] for k in names}
    else:
        import torch
        world, rank = dist.get_world_size(), dist.get_rank()
        all_counts = [torch.zeros(n_motifs, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(all_counts, torch.from_numpy(counts.copy()))
        parts_counts = np.stack([c.numpy() for c in all_counts])
        totals = parts_counts.sum(axis=1)
        width = int(totals.max())
        parts = {}
        for k in names:
            arr = np.ascontiguousarray(local[k])
            padded = np.zeros(width, dtype=arr.dtype)
            padded[:arr.size] = arr
            recv = [torch.zeros(width, dtype=torch.from_numpy(padded).dtype) for _ in range(world)] if rank == dst else None
            dist.gather(torch.from_numpy(padded), recv, dst=dst)
            if rank == dst:
                parts[k] = [recv[r].numpy()[:int(totals[r])] for r in range(world)]
        if rank != dst:
            return None
    out = {k: engine.merge_motif_major(parts_counts, parts[k]) for k in names}
    out["counts"] = parts_counts.sum(axis=0)
    out["motif"] = np.repeat(np.arange(n_motifs, dtype=np.int32), out["counts"])
    return out


def result_to_local(res, n_motifs, seq_global_ids):
    """engine.ScanResult -> the dict gather_sites expects (sequence ids mapped to global ids)."""
    return dict(counts=res.counts.copy(), seq=np.asarray(seq_global_ids, dtype=np.int64)[res.seq_idx],
                start=res.start.astype(np.int64), score=res.score.copy(), strand=res.strand.copy())
