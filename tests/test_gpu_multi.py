"""Round-2 product paths on the GPU, each against the oracle bit for bit:

* packed upload (msb_seqs_from_packed / msb_seqs_to_packed) == ASCII upload;
* MSB_SCAN_COUNTS (per-motif counts without ordering the sites) == the counts of the full scan;
* MSB_SCAN_ASYNC (the copy of scan k's sites overlaps scan k + 1) == the synchronous results;
* the sharded genome scan (`GenomeScanner`: units, shares, several devices and / or several ranks, async
  upload and download) == the reference's scan of whole chromosomes;
* `Scanner(devices=[...])`: regions dealt to several devices and gathered == one device.

Multi-device cases use every visible GPU (2..8 on a multi-GPU box); on a one-GPU box the same code runs
with the device listed several times, which exercises the planning, the threads and the gather.
"""
import numpy as np
import pytest

import oracle
from motifscan_b200 import _lib, engine, synth
from motifscan_b200.genome import DeviceGenome, PackedGenome
from motifscan_b200.genome_scan import GenomeScanner, scan_genome
from motifscan_b200.scanner import Scanner
from test_gpu_parity import assert_scan_equal, cutoffs_for, synth_pwms, synth_seqs

pytestmark = pytest.mark.gpu


def device_lists():
    n = _lib.device_count()
    lists = [[0], [0, 0, 0]]
    if n >= 2:
        lists.append(list(range(min(n, 8))))
    return lists


def test_packed_upload_equals_ascii_upload():
    rng = np.random.default_rng(5)
    seqs = synth_seqs(rng, 300, 1, 400, p_n=0.01, n_blocks=True) + ["", "ACGT" * 8, "N" * 33, "acgtn"]
    ctx = engine.default_context(0)
    a = engine.SequenceSet(ctx, seqs)
    codes, nmask = a.to_packed()
    lens = np.diff(a.seq_off)
    assert nmask.size == int(((lens + 31) // 32).sum()) and codes.size == 2 * nmask.size
    # host restatement of the layout
    want_c = np.concatenate([synth.pack_ascii(np.frombuffer(s.encode(), dtype=np.uint8))[0] for s in seqs if s])
    want_m = np.concatenate([synth.pack_ascii(np.frombuffer(s.encode(), dtype=np.uint8))[1] for s in seqs if s])
    assert np.array_equal(codes, want_c) and np.array_equal(nmask, want_m)
    # garbage behind the last base of a sequence and under masked bases must be ignored
    dirty_c, dirty_m = codes.copy(), nmask.copy()
    blk = 0
    for n in lens:
        nb = (int(n) + 31) // 32
        if nb and n % 32:
            tail = int(n) % 32
            dirty_m[blk + nb - 1] |= np.uint32((0xFFFFFFFF << tail) & 0xFFFFFFFF)
            dirty_c[2 * (blk + nb - 1) + 1] |= np.uint32(0xFFFF0000) if tail <= 24 else np.uint32(0)
        blk += nb
    for async_ in (False, True):
        b = engine.SequenceSet.from_packed(ctx, lens, dirty_c, dirty_m, async_=async_)
        assert np.array_equal(b.codes(), a.codes())
        c2, m2 = b.to_packed()
        assert np.array_equal(c2, codes) and np.array_equal(m2, nmask)
        b.close()
    a.close()


def test_counts_only_and_async_scans():
    rng = np.random.default_rng(6)
    pwms = synth_pwms(rng, 80)
    ctx = engine.Context(0)
    jobs = []
    for k in range(5):
        seqs = synth_seqs(rng, 60 + 30 * k, 300, 1500, p_n=0.002, n_blocks=True)
        cutoffs = cutoffs_for(pwms, seqs[:50], 1e-3)
        jobs.append((seqs, cutoffs, oracle.scan_arrays(pwms, cutoffs, seqs, 3, n_threads=8)))
    motifs = engine.MotifSet(ctx, pwms)
    ssets = [engine.SequenceSet(ctx, j[0]) for j in jobs]
    # counts only
    for (seqs, cutoffs, expect), sset in zip(jobs, ssets):
        motifs.set_cutoffs(cutoffs)
        n = engine.scan_device(ctx, motifs, sset, 3, counts_only=True)
        assert n == int(expect[0].sum())
        assert np.array_equal(ctx.site_counts(len(pwms)), expect[0])
        with pytest.raises(ValueError):
            ctx.region_counts(len(pwms))
    # five asynchronous scans in flight, looked at afterwards in reverse order
    results = []
    for (seqs, cutoffs, expect), sset in zip(jobs, ssets):
        motifs.set_cutoffs(cutoffs)
        results.append(engine.scan(ctx, motifs, sset, 3, async_=True))
        assert results[-1].n_sites == int(expect[0].sum())
    for res, (_, _, expect) in reversed(list(zip(results, jobs))):
        assert_scan_equal(res, expect)
        res.close()
    # compact site records (12 bytes over PCIe): decoded on the host to the same arrays; raw form = position << 1 | strand
    for (seqs, cutoffs, expect), sset in zip(jobs[:2], ssets[:2]):
        motifs.set_cutoffs(cutoffs)
        res = engine.scan(ctx, motifs, sset, 3, async_=True, compact=True)
        ps, sc, poff = res.compact()
        assert len(poff) == len(seqs) + 1 and np.array_equal(sc.view(np.uint64), expect[3].view(np.uint64))
        owner = np.searchsorted(poff, ps >> 1, side="right") - 1
        assert np.array_equal(owner, expect[1]) and np.array_equal((ps >> 1) - poff[owner], expect[2])
        assert np.array_equal((ps & 1) + 1, expect[4])
        assert_scan_equal(res, expect)
        res.close()
    # de-duplication + async (+ compact, which de-duplication overrides)
    motifs.set_cutoffs(jobs[0][1])
    c = engine.scan(ctx, motifs, ssets[0], 3, remove_dup=True, compact=True)
    assert c.compact() is None
    c.close()
    a = engine.scan(ctx, motifs, ssets[0], 3, remove_dup=True, async_=True)
    b = engine.scan(ctx, motifs, ssets[0], 3, remove_dup=True)
    assert np.array_equal(a.start, b.start) and np.array_equal(a.seq_idx, b.seq_idx) and a.n_sites < jobs[0][2][0].sum()
    assert np.array_equal(a.score.view(np.uint64), b.score.view(np.uint64))
    a.close(), b.close()
    for s in ssets:
        s.close()
    motifs.close(), ctx.close()


@pytest.fixture(scope="module")
def toy():
    rng = np.random.default_rng(41)
    pg = synth.packed_genome([70001, 123457, 1571, 33, 40000], names=["chr2", "chr1", "chrM", "chr10", "chrX"], seed=41)
    pwms = synth_pwms(rng, 70)
    whole = [pg.decode_bytes(c, 0, pg.chrom_sizes[c]).decode() for c in pg.chroms]
    cutoffs = cutoffs_for(pwms, [w for w in whole if len(w) > 1000][:2], 4e-4)
    expect = oracle.scan_arrays(pwms, cutoffs, whole, 3, n_threads=8)
    return pg, pwms, cutoffs, expect


def assert_genome_equal(sites, expect):
    counts, seq_idx, start, score, strand = expect
    assert np.array_equal(sites.counts, counts)
    assert np.array_equal(sites.chrom_idx, seq_idx.astype(np.int32))
    assert np.array_equal(sites.start, start.astype(np.int32))
    assert np.array_equal(sites.strand, strand)
    assert np.array_equal(sites.score.view(np.uint64), score.view(np.uint64))


@pytest.mark.parametrize("unit_bp", [1 << 26, 50000, 4096, 64])
def test_sharded_genome_scan_equals_reference(toy, unit_bp):
    pg, pwms, cutoffs, expect = toy
    assert expect[0].sum() > 2000
    for devices in device_lists():
        if unit_bp == 64 and len(devices) > 1:
            continue
        gs = GenomeScanner(pg, pwms, cutoffs=cutoffs, devices=devices, unit_bp=unit_bp, compact=(unit_bp != 50000))
        try:
            sites = gs.scan()
            assert_genome_equal(sites, expect)
            only = gs.scan(collect_sites=False)
            assert np.array_equal(only.counts, expect[0]) and len(only.start) == 0
        finally:
            gs.close()


def test_sharded_genome_scan_over_ranks_and_resident_units(toy):
    """world = 3 processes' shares scanned one after the other here: the per-rank counts add up and the
    per-rank site lists, gathered in rank order, are the reference's."""
    pg, pwms, cutoffs, expect = toy
    world = 3
    counts = np.zeros(len(pwms), dtype=np.int64)
    parts = []
    for rank in range(world):
        gs = GenomeScanner(pg, pwms, cutoffs=cutoffs, devices=[0, 0], world=world, rank=rank, unit_bp=30000, resident=True)
        first = gs.scan()
        again = gs.scan()          # second pass reads the units kept in HBM
        assert np.array_equal(first.counts, again.counts) and np.array_equal(first.start, again.start)
        counts += first.counts
        parts.append(first)
        gs.close()
    assert np.array_equal(counts, expect[0])
    cnt = np.stack([p.counts for p in parts])
    chrom = engine.merge_motif_major(cnt, [p.chrom_idx for p in parts])
    start = engine.merge_motif_major(cnt, [p.start for p in parts])
    score = engine.merge_motif_major(cnt, [p.score for p in parts])
    assert np.array_equal(chrom, expect[1].astype(np.int32)) and np.array_equal(start, expect[2].astype(np.int32))
    assert np.array_equal(score.view(np.uint64), expect[3].view(np.uint64))


def test_scan_genome_devices_argument_packs_a_host_genome(toy):
    pg, pwms, cutoffs, expect = toy

    class Host:           # an ASCII host genome: packed on the device first
        chroms, chrom_sizes = pg.chroms, pg.chrom_sizes
        fetch_bytes = staticmethod(pg.decode_bytes)

    sites = scan_genome(Host, pwms, cutoffs=cutoffs, devices=device_lists()[-1])
    assert_genome_equal(sites, expect)
    packed = PackedGenome.from_genome(Host)
    assert np.array_equal(packed.codes, pg.codes) and np.array_equal(packed.nmask, pg.nmask)


def test_packed_genome_cache_roundtrip(toy, tmp_path):
    pg = toy[0]
    pg.save(str(tmp_path / "toy"))
    back = PackedGenome.load(str(tmp_path / "toy"))
    assert back.chroms == pg.chroms and back.chrom_sizes == pg.chrom_sizes
    assert np.array_equal(back.codes, pg.codes) and np.array_equal(back.nmask, pg.nmask)


class Region:
    def __init__(self, chrom, start, end):
        self.chrom, self.start, self.end = chrom, start, end
        self.summit = (start + end) // 2


class Pwm:
    def __init__(self, matrix, cutoff):
        self.matrix, self.length, self.cutoffs = matrix, len(matrix[0]), {"1e-4": cutoff}


@pytest.mark.parametrize("source", ["ascii", "packed"])
def test_scanner_on_several_devices_equals_reference(toy, source):
    pg, pwms, cutoffs, _ = toy
    rng = np.random.default_rng(43)
    regions = []
    for _ in range(400):
        c = pg.chroms[int(rng.integers(0, len(pg.chroms)))]
        a = int(rng.integers(0, max(pg.chrom_sizes[c] - 50, 1)))
        regions.append(Region(c, a, a + int(rng.integers(20, 900))))
    wrapped = [Pwm(m, c) for m, c in zip(pwms, cutoffs)]

    class Host:
        chroms, chrom_sizes = pg.chroms, pg.chrom_sizes
        fetch_bytes = staticmethod(pg.decode_bytes)
        fetch_sequence = staticmethod(lambda c, a, b: pg.decode_bytes(c, a, b).decode())

    genome = Host if source == "ascii" else pg
    seqs = [pg.decode_bytes(r.chrom, r.start, r.end).decode() for r in regions]
    for remove_dup in (False, True):
        want = oracle.make_motif_sites(oracle.c_scan_motif(pwms, cutoffs, seqs, 3), [r.start for r in regions])
        if remove_dup:
            want = oracle.deduplicate_motif_sites(want, [p.length for p in wrapped])
        for devices in device_lists():
            sc = Scanner(genome, regions, window_size=0, remove_dup=remove_dup, devices=devices)
            got = sc.scan_motifs(wrapped)
            assert got.n_sites().sum() > 500
            for m in range(0, len(pwms), 7):
                for r in range(0, len(regions), 11):
                    g = [(s.start, s.score, s.strand) for s in got[m][r]]
                    w = [(s.start, s.score, s.strand) for s in want[m][r]]
                    assert g == w, (m, r, devices)
            assert np.array_equal(got.n_sites(), [sum(len(c) for c in per) for per in want])
