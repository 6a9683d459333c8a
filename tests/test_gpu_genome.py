"""Genome-wide chunked scan (configs[3]) == the reference's scan of whole chromosomes, bit for bit:
neither the resident path (position ranges of chromosomes kept in HBM) nor the streamed path
(chunks with overlap + device-side start limits) may lose or duplicate a window, for any chunk
size and any rank count.  Also: windows cut out of the resident genome on the device
(msb_seqs_extract) == the same windows sent as ASCII."""
import numpy as np
import pytest

import oracle
from motifscan_b200 import engine
from motifscan_b200.genome import DeviceGenome
from motifscan_b200.genome_scan import plan_chunks, plan_ranges, scan_genome
from test_gpu_parity import cutoffs_for, synth_pwms

pytestmark = pytest.mark.gpu


class ToyGenome:
    def __init__(self, rng, sizes):
        alphabet = np.frombuffer(b"ACGTacgt", dtype=np.uint8)
        self.seqs = {}
        for name, n in sizes.items():
            s = alphabet[rng.choice(8, size=n, p=[.2, .15, .15, .2, .09, .06, .06, .09])].copy()
            for _ in range(4 if n > 1000 else 0):
                a = int(rng.integers(0, n - 300))
                s[a:a + int(rng.integers(1, 250))] = ord("N")
            s[:min(40, n // 2)] = ord("N")            # telomere-like run at the chromosome start
            self.seqs[name] = bytes(s)
        self.chroms = sorted(self.seqs)
        self.chrom_sizes = {c: len(self.seqs[c]) for c in self.chroms}

    def fetch_bytes(self, chrom, start, end):
        return self.seqs[chrom][start:end]

    def fetch_sequence(self, chrom, start, end):
        return self.seqs[chrom][start:end].decode()


@pytest.fixture(scope="module")
def setup():
    rng = np.random.default_rng(31)
    genome = ToyGenome(rng, {"chr2": 70001, "chr1": 123457, "chrM": 1571, "chr10": 33})
    pwms = synth_pwms(rng, 60)
    whole = [genome.seqs[c].decode() for c in genome.chroms]
    cutoffs = cutoffs_for(pwms, whole[:2], 4e-4)
    expect = oracle.scan_arrays(pwms, cutoffs, whole, 3, n_threads=8)
    return genome, pwms, cutoffs, expect


def assert_same(sites, expect):
    counts, seq_idx, start, score, strand = expect
    assert np.array_equal(sites.counts, counts)
    assert np.array_equal(sites.chrom_idx, seq_idx.astype(np.int32))
    assert np.array_equal(sites.start, start.astype(np.int64))
    assert np.array_equal(sites.strand, strand)
    assert np.array_equal(sites.score.view(np.uint64), score.view(np.uint64))


@pytest.mark.parametrize("resident", [True, False])
@pytest.mark.parametrize("chunk_bp,batch_bp", [(1 << 22, 1 << 28), (10000, 64000), (4099, 9000), (512, 5000)])
def test_chunked_scan_equals_whole_chromosome_scan(setup, chunk_bp, batch_bp, resident):
    genome, pwms, cutoffs, expect = setup
    sites = scan_genome(genome, pwms, cutoffs=cutoffs, chunk_bp=chunk_bp, batch_bp=batch_bp,
                        ctx=engine.default_context(0), resident=resident)
    assert_same(sites, expect)
    assert len(sites) > 1000


def test_rank_shards_partition_the_genome(setup):
    """Three ranks: chunk ownership is a partition, per-rank counts add up, the merged sites are the
    unsharded list (no exchange step: the merge is a concatenation + sort on the host)."""
    genome, pwms, cutoffs, expect = setup
    world = 3
    plans = [plan_chunks(genome.chrom_sizes, 7000, 29, world, r) for r in range(world)]
    all_chunks = sorted(c for p in plans for c in p)
    assert all_chunks == sorted(plan_chunks(genome.chrom_sizes, 7000, 29))
    parts = [scan_genome(genome, pwms, cutoffs=cutoffs, chunk_bp=7000, batch_bp=50000, world=world, rank=r,
                         ctx=engine.default_context(0), resident=False) for r in range(world)]
    assert np.array_equal(sum(p.counts for p in parts), expect[0])
    motif = np.concatenate([p.motif for p in parts])
    cidx = np.concatenate([p.chrom_idx for p in parts])
    start = np.concatenate([p.start for p in parts])
    score = np.concatenate([p.score for p in parts])
    strand = np.concatenate([p.strand for p in parts])
    order = np.lexsort((strand, start, cidx, motif))
    assert np.array_equal(cidx[order], expect[1].astype(np.int32))
    assert np.array_equal(start[order], expect[2].astype(np.int64))
    assert np.array_equal(score[order].view(np.uint64), expect[4 - 1].view(np.uint64))
    # loads are balanced by the longest-first deal
    loads = [sum(c[2] - c[1] for c in p) for p in plans]
    assert max(loads) - min(loads) <= 7000


def merged(parts):
    motif = np.concatenate([p.motif for p in parts])
    cidx = np.concatenate([p.chrom_idx for p in parts])
    start = np.concatenate([p.start for p in parts])
    score = np.concatenate([p.score for p in parts])
    strand = np.concatenate([p.strand for p in parts])
    order = np.lexsort((strand, start, cidx, motif))
    return motif[order], cidx[order], start[order], score[order], strand[order]


def test_resident_rank_shards_partition_the_genome(setup):
    """Resident path, three ranks sharing ONE device-resident genome: round-robin ranges are a
    partition, per-rank counts add up, the merged sites are the unsharded list."""
    genome, pwms, cutoffs, expect = setup
    world = 3
    plans = [plan_ranges(genome.chrom_sizes, 6000, world, r) for r in range(world)]
    assert sorted(c for p in plans for c in p) == sorted(plan_ranges(genome.chrom_sizes, 6000))
    dg = DeviceGenome(genome, engine.default_context(0))
    parts = [scan_genome(dg, pwms, cutoffs=cutoffs, chunk_bp=6000, batch_bp=40000, world=world, rank=r)
             for r in range(world)]
    dg.close()
    assert np.array_equal(sum(p.counts for p in parts), expect[0])
    _, cidx, start, score, strand = merged(parts)
    assert np.array_equal(cidx, expect[1].astype(np.int32))
    assert np.array_equal(start, expect[2].astype(np.int64))
    assert np.array_equal(score.view(np.uint64), expect[3].view(np.uint64))
    assert np.array_equal(strand, expect[4])


def test_scan_ranges_unaligned(setup):
    """msb_scan_ranges with arbitrary (not tile-aligned) ranges == the whole-chromosome sites whose
    start falls in a range; overlapping ranges are refused."""
    genome, pwms, cutoffs, expect = setup
    counts, seq_idx, start, score, strand = expect
    ctx = engine.default_context(0)
    dg = DeviceGenome(genome, ctx)
    motifs = engine.MotifSet(ctx, pwms, cutoffs)
    rng = np.random.default_rng(5)
    i1, i2 = dg.chrom_index["chr1"], dg.chrom_index["chr2"]
    ranges = [(i1, 0, 1), (i1, 37, 1000), (i1, 1000, 1531), (i1, 50001, 123457 + 50), (i2, 511, 513), (i2, 69990, 70001),
              (dg.chrom_index["chrM"], 3, 1571), (dg.chrom_index["chr10"], 0, 33), (i2, 600, 600)]
    res = engine.scan_ranges(ctx, motifs, dg.seqs, 3, ranges)
    motif_of = np.repeat(np.arange(len(pwms)), counts)
    keep = np.zeros(len(start), dtype=bool)
    for c, a, b in ranges:
        keep |= (seq_idx == c) & (start >= a) & (start < b)
    assert np.array_equal(res.counts, np.bincount(motif_of[keep], minlength=len(pwms)))
    assert np.array_equal(res.seq_idx, seq_idx[keep].astype(np.int32))
    assert np.array_equal(res.start, start[keep].astype(np.int32))
    assert np.array_equal(res.strand, strand[keep])
    assert np.array_equal(res.score.view(np.uint64), score[keep].view(np.uint64))
    assert keep.sum() > 500 and (~keep).sum() > 500
    res.close()
    with pytest.raises(ValueError):
        engine.scan_ranges(ctx, motifs, dg.seqs, 3, [(i1, 0, 100), (i1, 99, 200)])
    with pytest.raises(ValueError):
        engine.scan_ranges(ctx, motifs, dg.seqs, 3, [(99, 0, 100)])
    empty = engine.scan_ranges(ctx, motifs, dg.seqs, 3, [])
    assert empty.n_sites == 0
    empty.close(), motifs.close(), dg.close()


def test_extract_equals_ascii_windows(setup):
    """Windows cut out of the resident genome (any alignment, clipped ends, empty intervals) are the
    same packed sequences -- codes and scan results -- as the same windows sent as ASCII."""
    genome, pwms, cutoffs, _ = setup
    ctx = engine.default_context(0)
    dg = DeviceGenome(genome, ctx)
    rng = np.random.default_rng(9)
    chroms, starts, ends = [], [], []
    for _ in range(300):
        c = genome.chroms[int(rng.integers(len(genome.chroms)))]
        n = genome.chrom_sizes[c]
        a = int(rng.integers(0, n))
        b = a + int(rng.integers(0, 700))          # may run past the chromosome end: clipped
        chroms.append(c), starts.append(a), ends.append(b)
    chroms += ["chr1", "chr1", "chr10", "chr2"]
    starts += [0, 123457, 0, 70000]
    ends += [32, 123500, 33, 70001]
    want_seqs = [genome.fetch_sequence(c, a, b) for c, a, b in zip(chroms, starts, ends)]
    a_set = engine.SequenceSet(ctx, want_seqs)
    x_set = dg.extract(chroms, starts, ends)
    assert x_set.n == a_set.n and x_set.total_bp == a_set.total_bp
    assert np.array_equal(x_set.seq_off, a_set.seq_off)
    assert np.array_equal(x_set.codes(), a_set.codes())
    motifs = engine.MotifSet(ctx, pwms, cutoffs)
    ra, rx = engine.scan(ctx, motifs, a_set, 3), engine.scan(ctx, motifs, x_set, 3)
    assert ra.n_sites > 50
    assert np.array_equal(ra.counts, rx.counts) and np.array_equal(ra.seq_idx, rx.seq_idx)
    assert np.array_equal(ra.start, rx.start) and np.array_equal(ra.strand, rx.strand)
    assert np.array_equal(ra.score.view(np.uint64), rx.score.view(np.uint64))
    with pytest.raises(KeyError):
        dg.extract(["chr9"], [0], [10])
    with pytest.raises(ValueError):
        dg.seqs.extract([0], [-1], [10])
    ra.close(), rx.close(), motifs.close(), a_set.close(), x_set.close(), dg.close()


def test_counts_only_mode(setup):
    genome, pwms, cutoffs, expect = setup
    sites = scan_genome(genome, pwms, cutoffs=cutoffs, chunk_bp=20000, batch_bp=100000, collect_sites=False,
                        ctx=engine.default_context(0))
    assert np.array_equal(sites.counts, expect[0])
    assert sites.start.size == 0


def test_window_ncount_and_empty_extract(setup):
    """msb_seqs_window_ncount == counting the non-ACGT letters of the fetched window (clipped at the
    chromosome end); extracting nothing gives an empty set that scans to nothing."""
    genome, pwms, cutoffs, _ = setup
    ctx = engine.default_context(0)
    dg = DeviceGenome(genome, ctx)
    rng = np.random.default_rng(12)
    idx = rng.integers(0, len(genome.chroms), size=500).astype(np.int32)
    starts = np.array([rng.integers(0, genome.chrom_sizes[genome.chroms[i]]) for i in idx], dtype=np.int64)
    for length in (1, 30, 33, 700):
        got = dg.seqs.window_ncount(idx, starts, length)
        want = [sum(ch not in "ACGTacgt" for ch in genome.fetch_sequence(genome.chroms[i], int(a), int(a) + length))
                for i, a in zip(idx, starts)]
        assert got.tolist() == want
    assert sum(want) > 0
    empty = dg.seqs.extract(np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.int64))
    assert empty.n == 0 and empty.total_bp == 0
    motifs = engine.MotifSet(ctx, pwms, cutoffs)
    res = engine.scan(ctx, motifs, empty, 3)
    assert res.n_sites == 0
    res.close(), motifs.close(), empty.close(), dg.close()


def test_range_scan_with_several_prefilter_batches(setup):
    """1,900 motifs need several prefilter launches per range (24 K-step units of B each); ranges and
    batches multiply, the candidate records of all launches accumulate in the same lane buffers."""
    genome, _, _, _ = setup
    rng = np.random.default_rng(77)
    pwms = synth_pwms(rng, 1900)
    whole = [genome.seqs[c].decode() for c in genome.chroms]
    cutoffs = cutoffs_for(pwms, whole[:2], 2e-4)
    ctx = engine.default_context(0)
    dg = DeviceGenome(genome, ctx)
    motifs = engine.MotifSet(ctx, pwms, cutoffs)
    full = engine.scan(ctx, motifs, dg.seqs, 3)
    n_batches = ctx.counters()["prefilter_launches"]
    assert n_batches >= 2
    i1 = dg.chrom_index["chr1"]
    ranges = [(i1, 0, 40000), (i1, 40000, 90001), (i1, 90001, 123457), (dg.chrom_index["chr2"], 0, 70001),
              (dg.chrom_index["chrM"], 0, 1571), (dg.chrom_index["chr10"], 0, 33)]
    parts = engine.scan_ranges(ctx, motifs, dg.seqs, 3, ranges)
    assert ctx.counters()["prefilter_launches"] == n_batches * len(ranges)
    assert parts.n_sites == full.n_sites > 10000
    assert np.array_equal(parts.counts, full.counts) and np.array_equal(parts.seq_idx, full.seq_idx)
    assert np.array_equal(parts.start, full.start) and np.array_equal(parts.strand, full.strand)
    assert np.array_equal(parts.score.view(np.uint64), full.score.view(np.uint64))
    # and against the oracle for a handful of motifs
    pick = [0, 7, 950, 1899]
    want = oracle.scan_arrays([pwms[m] for m in pick], [cutoffs[m] for m in pick], whole, 3, n_threads=8)
    for k, m in enumerate(pick):
        a, b = full.offsets[m], full.offsets[m + 1]
        wa = int(want[0][:k].sum())
        wb = wa + int(want[0][k])
        assert np.array_equal(full.seq_idx[a:b], want[1][wa:wb].astype(np.int32))
        assert np.array_equal(full.start[a:b], want[2][wa:wb].astype(np.int32))
        assert np.array_equal(full.score[a:b].view(np.uint64), want[3][wa:wb].view(np.uint64))
    parts.close(), full.close(), motifs.close(), dg.close()
