"""CPU tests of the host logic added for the multi-GPU paths: share / unit planning over the packed genome,
the packed layout's numpy restatement, the host-side gather (msb_merge_motif_major is plain host code in the
library: it loads and runs without a GPU), and the world-size-2 count gather over gloo."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_merge_motif_major_equals_a_stable_sort():
    from motifscan_b200 import engine
    rng = np.random.default_rng(1)
    for n_parts, n_motifs, scale in ((1, 5, 10), (4, 750, 40), (8, 13, 200000), (3, 1, 7), (5, 40, 0)):
        counts = rng.integers(0, scale + 1, size=(n_parts, n_motifs)).astype(np.int64)
        parts_motif = [np.repeat(np.arange(n_motifs), counts[p]) for p in range(n_parts)]
        for dtype in (np.int32, np.float64, np.int8):
            arrays = [rng.integers(-100, 100, size=int(counts[p].sum())).astype(dtype) for p in range(n_parts)]
            got = engine.merge_motif_major(counts, arrays)
            motif = np.concatenate(parts_motif)
            order = np.argsort(motif, kind="stable")          # part order is kept within a motif
            assert np.array_equal(got, np.concatenate(arrays)[order])
        seq = [rng.integers(0, 1000, size=int(counts[p].sum())).astype(np.int32) for p in range(n_parts)]
        add = np.arange(n_parts, dtype=np.int64) * 1000
        got = engine.merge_motif_major(counts, seq, add=add)
        want = np.concatenate([s + int(a) for s, a in zip(seq, add)])[np.argsort(np.concatenate(parts_motif), kind="stable")]
        assert np.array_equal(got, want)
    assert engine.merge_motif_major(np.zeros((0, 4), dtype=np.int64), []).size == 0


def test_plan_units_partitions_hg19():
    from motifscan_b200.genome_scan import plan_units
    from motifscan_b200.synth import HG19_NAMES, HG19_SIZES
    order = sorted(range(len(HG19_NAMES)), key=lambda i: HG19_NAMES[i])
    sizes = [HG19_SIZES[i] for i in order]
    block_off = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum([(s + 31) // 32 for s in sizes], out=block_off[1:])
    for n_shares in (1, 2, 4, 8):
        shares = plan_units(block_off, sizes, n_shares, (1 << 26) // 32, 29)
        assert len(shares) == n_shares
        per_share = [sum(u.owned_bp for u in units) for units in shares]
        assert sum(per_share) == sum(sizes)
        assert max(per_share) / (sum(per_share) / n_shares) < 1.001           # shares balance to 0.1 %
        owned = {c: [] for c in range(len(sizes))}
        for units in shares:
            for u in units:
                assert u.block1 - u.block0 <= (1 << 26) // 32 and u.upload1 - u.block1 <= 1
                at = u.block0
                for c, a, b, f in u.pieces:
                    assert a % 32 == 0 and a < b <= f <= sizes[c] and f - b == min(32, sizes[c] - b)
                    assert block_off[c] + a // 32 == at       # the pieces tile the uploaded planes
                    at += (f - a + 31) // 32
                    owned[c].append((a, b))
                assert at == u.upload1
        for c, spans in owned.items():                         # every base owned exactly once
            spans.sort()
            assert spans[0][0] == 0 and spans[-1][1] == sizes[c]
            assert all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))


def test_plan_units_balanced_by_work():
    """Shares cut at equal work (non-N bases) still own every base exactly once, and even out the work."""
    from motifscan_b200 import synth
    from motifscan_b200.genome_scan import plan_units
    pg = synth.packed_genome([200000, 120000, 50000, 90000, 40], seed=4)
    sizes = [pg.chrom_sizes[c] for c in pg.chroms]
    w = pg.work_prefix()
    acgt = sum(sizes) - sum(pg.decode_bytes(c, 0, 1 << 40).count(b"N") for c in pg.chroms)
    assert 0 <= w[-1] - (acgt + pg.n_blocks) < 32 * len(sizes)     # the padding behind a chromosome's last base counts as bases
    for n in (1, 2, 3, 8):
        shares = plan_units(pg.block_off, sizes, n, 500, 29, work_prefix=w)
        assert sum(u.owned_bp for units in shares for u in units) == sum(sizes)
        flat = [u for units in shares for u in units]
        assert flat[0].block0 == 0 and flat[-1].block1 == pg.n_blocks
        assert all(a.block1 == b.block0 for a, b in zip(flat, flat[1:]))
        work = np.array([sum(w[u.block1] - w[u.block0] for u in units) for units in shares], dtype=float)
        by_size = plan_units(pg.block_off, sizes, n, 500, 29)
        work_by_size = np.array([sum(w[u.block1] - w[u.block0] for u in units) for units in by_size], dtype=float)
        assert work.max() / work.mean() < 1.002 and work.max() <= work_by_size.max() + 1


def test_packed_layout_restatement_and_decode():
    from motifscan_b200 import synth
    import oracle
    rng = np.random.default_rng(2)
    alphabet = np.frombuffer(b"ACGTacgtNnRy-", dtype=np.uint8)
    for n in (1, 15, 16, 17, 31, 32, 33, 1000):
        s = alphabet[rng.integers(0, len(alphabet), size=n)]
        codes, nmask = synth.pack_ascii(s)
        assert nmask.size == (n + 31) // 32 and codes.size == 2 * nmask.size
        want = oracle.encode(bytes(s).decode())               # the reference's int8 codes (cscore.c:81-114)
        two = ((codes[:, None] >> (2 * np.arange(16, dtype=np.uint32))[None, :]) & 3).ravel()[:n].astype(np.int8)
        isn = ((nmask[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).ravel()[:n].astype(bool)
        assert np.array_equal(np.where(isn, -1, two), want)
        assert not two[isn].any()                             # N -> code 0
    pg = synth.packed_genome([5000, 333, 64, 10001, 31], seed=3)
    for c in pg.chroms:
        whole = pg.decode_bytes(c, 0, 1 << 40)
        assert len(whole) == pg.chrom_sizes[c]
        for a, b in ((0, 1), (5, 77), (31, 33), (100, 100), (pg.chrom_sizes[c] - 3, pg.chrom_sizes[c] + 9)):
            assert pg.decode_bytes(c, a, b) == whole[max(a, 0):b]
    with pytest.raises(KeyError):
        pg.decode_bytes("chrNope", 0, 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import oracle
    from motifscan_b200 import shard, synth
    from motifscan_b200.genome_scan import plan_units
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # a sharded genome scan with the oracle as the per-unit scanner: this rank's units -> counts -> gather
        pg = synth.packed_genome([9000, 4001, 700], seed=8)
        rng = np.random.default_rng(9)
        pwms = [np.around(rng.normal(0, 2, size=(4, int(rng.integers(4, 20)))), 5).tolist() for _ in range(6)]
        cutoffs = [0.4] * len(pwms)
        sizes = [pg.chrom_sizes[c] for c in pg.chroms]
        shares = plan_units(pg.block_off, sizes, world, 40, max(len(p[0]) for p in pwms) - 1)
        counts = np.zeros(len(pwms), dtype=np.int64)
        for unit in shares[rank]:
            for c, a, b, f in unit.pieces:
                seq = pg.decode_bytes(pg.chroms[c], a, f).decode()
                cnt, _, start, _, _ = oracle.scan_arrays(pwms, cutoffs, [seq], 3)
                motif = np.repeat(np.arange(len(pwms)), cnt)
                counts += np.bincount(motif[start < b - a], minlength=len(pwms))    # starts in the halo belong to the next unit
        total = shard.gather_counts(counts, dist)                 # all-reduce form: every rank gets the sum
        at_root = shard.gather_counts(counts, dist, dst=0)        # one-hop form: rank 0 gets the sum, the others None
        whole = [pg.decode_bytes(c, 0, pg.chrom_sizes[c]).decode() for c in pg.chroms]
        full = oracle.scan_arrays(pwms, cutoffs, whole, 3)[0]
        root_ok = np.array_equal(at_root, full) if rank == 0 else at_root is None
        q.put(bool(np.array_equal(total, full) and root_ok and full.sum() > 100 and counts.sum() < full.sum()))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_two_rank_sharded_count_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True and q.get(timeout=5) is True


def test_merge_sites_translates_local_sequence_numbers():
    from motifscan_b200 import engine
    rng = np.random.default_rng(3)
    for n_parts, n_motifs, scale in ((1, 4, 20), (5, 60, 3000), (3, 750, 50)):
        counts = rng.integers(0, scale + 1, size=(n_parts, n_motifs)).astype(np.int64)
        n_seq = [int(rng.integers(1, 9)) for _ in range(n_parts)]
        seq = [rng.integers(0, n_seq[p], size=int(counts[p].sum())).astype(np.int32) for p in range(n_parts)]
        start = [rng.integers(0, 10**6, size=len(s)).astype(np.int32) for s in seq]
        score = [rng.normal(size=len(s)) for s in seq]
        strand = [rng.integers(1, 3, size=len(s)).astype(np.int8) for s in seq]
        group = [rng.integers(0, 25, size=n).astype(np.int32) for n in n_seq]
        offset = [rng.integers(0, 10**8, size=n).astype(np.int32) for n in n_seq]
        order = np.argsort(np.concatenate([np.repeat(np.arange(n_motifs), counts[p]) for p in range(n_parts)]), kind="stable")
        g, st, sc, sd = engine.merge_sites(counts, seq, start, score, strand, seq_to_group=group, seq_offset=offset)
        assert np.array_equal(g, np.concatenate([group[p][seq[p]] for p in range(n_parts)])[order])
        assert np.array_equal(st, np.concatenate([start[p] + offset[p][seq[p]] for p in range(n_parts)])[order])
        assert np.array_equal(sc, np.concatenate(score)[order]) and np.array_equal(sd, np.concatenate(strand)[order])
        g2, st2, _, _ = engine.merge_sites(counts, seq, start, score, strand)
        assert np.array_equal(g2, np.concatenate(seq)[order]) and np.array_equal(st2, np.concatenate(start)[order])


def test_native_site_tables_equal_the_python_writer(tmp_path):
    """msb_format_site_tables (host code in the library) writes motif_sites_number.xls / motif_sites_score.xls byte
    for byte like the list-based writer (io/__init__.py:12-38 restated): counts, `NA`, and doubles printed as
    Python's str() prints them -- fixed / scientific switch, shortest round-trip digits, denormals, huge values."""
    import struct
    from motifscan_b200 import io as msio
    from motifscan_b200.scanner import MotifSites
    rng = np.random.default_rng(0)
    n_pwms, n_regions = 23, 3000

    class Region:
        def __init__(self, c, a, b):
            self.chrom, self.start, self.end = c, a, b

    class Pwm:
        def __init__(self, i):
            self.matrix_id, self.name, self.length = f"MA{i:04d}.1", f"name{i}", 8
    regions = [Region(f"chr{i % 5}", i * 10, i * 10 + 500) for i in range(n_regions)]
    pwms = [Pwm(i) for i in range(n_pwms)]
    counts = rng.integers(0, 2500, size=n_pwms)
    counts[3] = 0
    seqs = [np.sort(rng.integers(0, n_regions, size=c)).astype(np.int32) for c in counts]
    vals = np.concatenate([rng.normal(0.5, 0.3, size=c) for c in counts])
    special = [0.0001, 0.00001, 1e16, 1e15, 1.0, -0.0, 123456789.125, 1e-7, 0.1 + 0.2, 5e-324, 1.7976931348623157e308,
               2.5, 100.0, 1e22, 1.5e-5, 0.3, 1 / 3, 2 / 3, 9.999999999999999e-05, 0.00010000000000000002]
    vals[:len(special)] = special
    bits = rng.integers(0, 2**63, size=2000, dtype=np.int64)        # arbitrary finite bit patterns
    rnd = np.array([struct.unpack("<d", struct.pack("<q", int(b)))[0] for b in bits])
    rnd = rnd[np.isfinite(rnd)]
    vals[len(special):len(special) + len(rnd)] = rnd

    class Res:
        pass
    res = Res()
    res.offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    res.seq_idx, res.start = np.concatenate(seqs), np.zeros(len(vals), np.int32)
    res.score, res.strand = vals, np.ones(len(vals), np.int8)
    sites = MotifSites(res, n_pwms, [g.start for g in regions], [8] * n_pwms)
    msio.write_sites_table(str(tmp_path / "native"), pwms, regions, sites)
    msio.write_sites_table(str(tmp_path / "lists"), pwms, regions, sites.tolist())
    for name in ("motif_sites_number.xls", "motif_sites_score.xls"):
        assert (tmp_path / "native" / name).read_bytes() == (tmp_path / "lists" / name).read_bytes(), name
