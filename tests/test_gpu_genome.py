"""Genome-wide chunked scan (configs[3]) == the reference's scan of whole chromosomes, bit for bit:
chunking with overlap + device-side start limits must neither lose nor duplicate a window, for any
chunk size and any rank count."""
import numpy as np
import pytest

import oracle
from motifscan_b200 import engine
from motifscan_b200.genome_scan import plan_chunks, scan_genome
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


@pytest.mark.parametrize("chunk_bp,batch_bp", [(1 << 22, 1 << 28), (10000, 64000), (4099, 9000), (512, 5000)])
def test_chunked_scan_equals_whole_chromosome_scan(setup, chunk_bp, batch_bp):
    genome, pwms, cutoffs, expect = setup
    sites = scan_genome(genome, pwms, cutoffs=cutoffs, chunk_bp=chunk_bp, batch_bp=batch_bp,
                        ctx=engine.default_context(0))
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
                         ctx=engine.default_context(0)) for r in range(world)]
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


def test_counts_only_mode(setup):
    genome, pwms, cutoffs, expect = setup
    sites = scan_genome(genome, pwms, cutoffs=cutoffs, chunk_bp=20000, batch_bp=100000, collect_sites=False,
                        ctx=engine.default_context(0))
    assert np.array_equal(sites.counts, expect[0])
    assert sites.start.size == 0
