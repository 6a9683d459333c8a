"""configs[1] at BASELINE.json's full size (750 motifs x 50,000 x 1 kb, both strands, p = 1e-4 cutoffs
built on the device) checked through size-independent properties, plus a random sample against the
oracle.  Everything bit-exact."""
import numpy as np
import pytest

import oracle
from motifscan_b200 import engine, synth

pytestmark = pytest.mark.gpu

N_MOTIFS, N_REGIONS, REGION_BP = 750, 50000, 1000


@pytest.fixture(scope="module")
def full():
    ctx = engine.default_context(0)
    _, pwms, _ = synth.motif_set(N_MOTIFS, seed=2020)
    blob, off = synth.peak_set(N_REGIONS, REGION_BP, seed=50)
    motifs = engine.MotifSet(ctx, pwms)
    lmax = max(p.shape[1] for p in pwms)
    bblob, boff = synth.background_samples(100000, lmax, seed=1)
    bg = engine.SequenceSet(ctx, blob=bblob, seq_off=boff)
    cutoffs = np.around(engine.score_select(ctx, motifs, bg, 3, [int(100000 * 1e-4) - 1])[:, 0], 8)
    bg.close()
    motifs.set_cutoffs(cutoffs)
    sset = engine.SequenceSet(ctx, blob=blob, seq_off=off)
    res = engine.scan(ctx, motifs, sset, 3).detach()
    yield ctx, pwms, cutoffs, blob, off, motifs, sset, res
    sset.close(), motifs.close()


def as_tuple(res):
    return res.counts, res.seq_idx, res.start, res.strand, res.score.view(np.uint64)


def test_order_and_bounds(full):
    _, pwms, cutoffs, _, _, _, _, res = full
    assert res.n_sites > 3_000_000
    lens = np.array([p.shape[1] for p in pwms])
    motif = np.repeat(np.arange(N_MOTIFS), res.counts)
    # reference order: motif, sequence, start, forward before reverse (cscore.c:336-389)
    key = ((motif.astype(np.int64) * N_REGIONS + res.seq_idx) * 2048 + res.start) * 2 + (res.strand - 1)
    assert np.all(np.diff(key) > 0)
    assert np.all(res.start >= 0) and np.all(res.start + lens[motif] <= REGION_BP)
    assert np.all(res.score - cutoffs[motif] >= -1e-10)          # the reference's predicate holds for every site


def test_partition_of_the_regions(full):
    """Scanning the regions in 3 uneven blocks gives the same sites as one scan (the sharding rule of
    SURVEY 8e: sequences are independent)."""
    ctx, _, _, blob, off, motifs, _, res = full
    cuts = [0, 7001, 31999, N_REGIONS]
    parts = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        s = engine.SequenceSet(ctx, blob=blob[off[a]:off[b]], seq_off=off[a:b + 1] - off[a])
        parts.append((a, engine.scan(ctx, motifs, s, 3).detach()))
        s.close()
    assert np.array_equal(sum(p.counts for _, p in parts), res.counts)
    for m in (0, 1, 373, 749):
        got_seq = np.concatenate([p.seq_idx[p.offsets[m]:p.offsets[m + 1]] + a for a, p in parts])
        got_start = np.concatenate([p.start[p.offsets[m]:p.offsets[m + 1]] for _, p in parts])
        got_score = np.concatenate([p.score[p.offsets[m]:p.offsets[m + 1]] for _, p in parts])
        sl = slice(res.offsets[m], res.offsets[m + 1])
        assert np.array_equal(got_seq, res.seq_idx[sl]) and np.array_equal(got_start, res.start[sl])
        assert np.array_equal(got_score.view(np.uint64), res.score[sl].view(np.uint64))


def test_strands_decompose(full):
    ctx, _, _, _, _, motifs, sset, res = full
    fwd = engine.scan(ctx, motifs, sset, 1).detach()
    rev = engine.scan(ctx, motifs, sset, 2).detach()
    assert np.array_equal(fwd.counts + rev.counts, res.counts)
    assert np.all(fwd.strand == 1) and np.all(rev.strand == 2)
    pick = res.strand == 1
    assert np.array_equal(res.seq_idx[pick], fwd.seq_idx) and np.array_equal(res.start[pick], fwd.start)
    assert np.array_equal(res.score[pick].view(np.uint64), fwd.score.view(np.uint64))
    assert np.array_equal(res.seq_idx[~pick], rev.seq_idx) and np.array_equal(res.start[~pick], rev.start)
    assert np.array_equal(res.score[~pick].view(np.uint64), rev.score.view(np.uint64))


def test_stricter_cutoffs_select_a_subset(full):
    """Raising every cutoff keeps exactly the sites that still satisfy score - cutoff >= -1e-10."""
    ctx, _, cutoffs, _, _, motifs, sset, res = full
    stricter = cutoffs + 0.03
    motifs.set_cutoffs(stricter)
    try:
        got = engine.scan(ctx, motifs, sset, 3).detach()
    finally:
        motifs.set_cutoffs(cutoffs)
    motif = np.repeat(np.arange(N_MOTIFS), res.counts)
    keep = res.score - stricter[motif] >= -1e-10
    assert 0 < keep.sum() < res.n_sites
    assert np.array_equal(got.counts, np.bincount(motif[keep], minlength=N_MOTIFS))
    assert np.array_equal(got.seq_idx, res.seq_idx[keep]) and np.array_equal(got.start, res.start[keep])
    assert np.array_equal(got.score.view(np.uint64), res.score[keep].view(np.uint64))


def test_scan_ascii_and_resident_extraction_agree(full):
    """The three ways in -- encoded set, one-call ASCII with sliced upload, windows cut out of a
    resident copy on the device -- give the same arrays."""
    ctx, _, _, blob, off, motifs, sset, res = full
    a = engine.scan_ascii(ctx, motifs, blob, off, 3).detach()
    for x, y in zip(as_tuple(a), as_tuple(res)):
        assert np.array_equal(x, y)
    # resident: the 50 Mbp of peaks as one "chromosome", every peak cut back out of it
    whole = engine.SequenceSet(ctx, blob=blob, seq_off=np.array([0, blob.size], dtype=np.int64))
    cut = whole.extract(np.zeros(N_REGIONS, dtype=np.int32), off[:-1], off[1:])
    r = engine.scan(ctx, motifs, cut, 3).detach()
    for x, y in zip(as_tuple(r), as_tuple(res)):
        assert np.array_equal(x, y)
    cut.close(), whole.close()


def test_random_sample_against_the_oracle(full):
    _, pwms, cutoffs, blob, off, _, _, res = full
    rng = np.random.default_rng(4)
    pick = np.sort(rng.choice(N_REGIONS, size=120, replace=False))
    raw = blob.tobytes()
    seqs = [raw[off[i]:off[i + 1]].decode() for i in pick]
    counts, seq_idx, start, score, strand = oracle.scan_arrays([p.tolist() for p in pwms], cutoffs.tolist(), seqs, 3, n_threads=8)
    motif = np.repeat(np.arange(N_MOTIFS), res.counts)
    sel = np.isin(res.seq_idx, pick)
    assert np.array_equal(np.bincount(motif[sel], minlength=N_MOTIFS), counts)
    assert np.array_equal(np.searchsorted(pick, res.seq_idx[sel]), seq_idx)
    assert np.array_equal(res.start[sel], start) and np.array_equal(res.strand[sel], strand)
    assert np.array_equal(res.score[sel].view(np.uint64), score.view(np.uint64))
    assert counts.sum() > 5000
