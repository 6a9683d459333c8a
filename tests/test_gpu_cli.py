"""End-to-end GPU test of the two CLI entry points (`motif --build`, `scan`) on a synthetic genome:
every output file must equal, byte for byte, what the reference's pipeline produces when its
native extension is replaced by the oracle (oracle.c restates cscore.c; the Python steps are the
reference's rules restated in oracle/__init__.py and motifscan_b200's writers, which
tests/test_host_widening.py pins to the reference's own output)."""
import os

import numpy as np
import pytest

import oracle
from motifscan_b200 import cli
from motifscan_b200 import io as msio
from motifscan_b200 import stats as msstats
from motifscan_b200.genome import Genome
from motifscan_b200.motif import MotifPwms, PositionWeightMatrix, pfm_to_pwm, read_jaspar_pfms
from motifscan_b200.region import generate_control_regions, load_motifscan_regions
from motifscan_b200.scanner import MotifSite, Scanner

pytestmark = pytest.mark.gpu

BG = dict(A=0.29, C=0.21, G=0.21, T=0.29)


@pytest.fixture(scope="module")
def workspace(tmp_path_factory):
    root = tmp_path_factory.mktemp("cli")
    rng = np.random.default_rng(5)
    gdir, mdir = root / "toyg", root / "toym"
    gdir.mkdir(), mdir.mkdir()
    alphabet = np.frombuffer(b"ACGTacgt", dtype=np.uint8)
    with open(gdir / "toyg.fa", "w") as fh:
        for chrom, n in [("chr1", 60000), ("chr2", 45000), ("chrX", 30011)]:
            seq = alphabet[rng.choice(8, size=n, p=[.2, .15, .15, .2, .09, .06, .06, .09])].copy()
            for _ in range(6):                      # N blocks, some short enough to sit inside windows
                a = int(rng.integers(0, n - 400))
                seq[a:a + int(rng.integers(1, 300))] = ord("N")
            text = bytes(seq).decode()
            fh.write(f">{chrom}\n")
            for i in range(0, n, 60):
                fh.write(text[i:i + 60] + "\n")
    with open(gdir / "toyg_bg_freq.txt", "w") as fh:
        fh.write("Base\tFrequency\n" + "".join(f"{b}\t{v}\n" for b, v in BG.items()))
    with open(mdir / "toym_pfms.jaspar", "w") as fh:
        for k in range(24):
            L = int(rng.integers(6, 22))
            pfm = np.round(int(rng.integers(20, 2000)) * rng.dirichlet([0.3] * 4, size=L).T).astype(int)
            pfm[0, pfm.sum(axis=0) == 0] = 1
            fh.write(f">MA{k:04d}.1\tTF{k}\n")
            for base, row in zip("ACGT", pfm):
                fh.write(f"{base}  [ " + " ".join(f"{v:5d}" for v in row) + " ]\n")
    with open(root / "peaks.bed", "w") as fh:
        sizes = {"chr1": 60000, "chr2": 45000, "chrX": 30011}
        for i in range(150):
            chrom = ["chr1", "chr2", "chrX"][int(rng.integers(3))]
            a = int(rng.integers(0, sizes[chrom] - 900))
            fh.write(f"{chrom}\t{a}\t{a + int(rng.integers(150, 900))}\tpeak{i}\t{float(rng.uniform(1, 50)):.2f}\n")
    return root


def oracle_build(gdir, mdir, n_random, seed):
    genome = Genome("toyg", path=str(gdir))
    pwms = MotifPwms(name="toym", genome="toyg")
    for matrix_id, name, pfm in read_jaspar_pfms(str(mdir / "toym_pfms.jaspar")):
        pwms.append(PositionWeightMatrix(pfm_to_pwm(pfm, genome.bg_freq), name=name, matrix_id=matrix_id))
    lmax = max(p.length for p in pwms)
    seqs = list(genome.random_sequences(n_random, lmax, 0, seed))
    scores = oracle.score_arrays([p.matrix for p in pwms], seqs, 3)
    for pwm, cut in zip(pwms, oracle.average_cutoffs([oracle.get_score_cutoffs(scores)])):
        for p, v in cut.items():
            pwm.set_cutoff(p, v)
    return genome, pwms


def oracle_sites(genome, regions, pwms, window, strand, p_value):
    sc = Scanner(genome=genome, regions=regions, window_size=window, strand=strand, p_value=p_value)
    flat = oracle.c_scan_motif([p.matrix.tolist() for p in pwms], [p.cutoffs[p_value] for p in pwms],
                               sc.sequences, {"+": 1, "-": 2, "both": 3}[strand], 4)
    nested = oracle.make_motif_sites(flat, sc.seq_starts)
    nested = oracle.deduplicate_motif_sites(nested, [p.length for p in pwms])
    return [[[MotifSite(*s) for s in cell] for cell in per] for per in nested]


def test_build_then_scan_files_identical(workspace, tmp_path):
    gdir, mdir = workspace / "toyg", workspace / "toym"
    assert cli.main(["motif", "--build", str(mdir), "-g", str(gdir), "--n-random", "20000", "--seed", "1"]) == 0
    built = mdir / "toym_toyg_pwms.motifscan"
    genome, pwms = oracle_build(gdir, mdir, 20000, 1)
    want = tmp_path / "want.motifscan"
    pwms.write_motifscan_pwms(str(want))
    assert built.read_text() == want.read_text()

    out = tmp_path / "out"
    assert cli.main(["scan", "-i", str(workspace / "peaks.bed"), "-m", str(mdir), "-g", str(gdir), "-o", str(out),
                     "-p", "1e-3", "-w", "400", "--site", "--n-random", "2", "--seed", "3"]) == 0
    built_pwms = MotifPwms(name="toym", genome="toyg")
    built_pwms.read_motifscan_pwms(str(built))
    regions = load_motifscan_regions(str(workspace / "peaks.bed"), "bed")
    controls = generate_control_regions(2, regions, genome.chrom_sizes, random_seed=3)
    sites = oracle_sites(genome, regions, built_pwms, 400, "both", "1e-3")
    ctl_sites = oracle_sites(genome, controls, built_pwms, 400, "both", "1e-3")
    ref = tmp_path / "ref"
    msio.write_sites_table(str(ref), built_pwms, regions, sites)
    msio.write_sites_bed(str(ref), built_pwms, regions, sites)
    msio.write_enrich_table(str(ref), msstats.motif_enrichment(built_pwms, sites, ctl_sites))
    n_files = 0
    for root, _, names in os.walk(ref):
        for name in names:
            rel = os.path.relpath(os.path.join(root, name), ref)
            assert (out / rel).read_text() == (ref / rel).read_text(), rel
            n_files += 1
    assert n_files == 3 + len(built_pwms)
    assert sum(len(c) for per in sites for c in per) > 100   # the comparison is not vacuous


def _tree(path):
    out = {}
    for root, _, names in os.walk(path):
        for name in names:
            full = os.path.join(root, name)
            out[os.path.relpath(full, path)] = open(full, "rb").read()
    return out


def test_scan_through_the_packed_genome_and_on_several_gpus(workspace, tmp_path, monkeypatch):
    """`scan` over the packed, device-resident genome (cache file next to the FASTA, windows cut on the device,
    tables written by the native formatter) and `--gpus N` write byte for byte the files of the plain run that
    fetches region strings from the FASTA like the reference."""
    from motifscan_b200 import _lib
    gdir, mdir = workspace / "toyg", workspace / "toym"
    if not (mdir / "toym_toyg_pwms.motifscan").exists():
        assert cli.main(["motif", "--build", str(mdir), "-g", str(gdir), "--n-random", "20000", "--seed", "1"]) == 0
    for stale in gdir.glob("toyg.packed.*"):
        stale.unlink()
    base = ["scan", "-i", str(workspace / "peaks.bed"), "-m", str(mdir), "-g", str(gdir), "-p", "1e-3", "-w", "400",
            "--site", "--n-random", "2", "--seed", "3"]
    assert cli.main(base + ["-o", str(tmp_path / "plain")]) == 0
    assert not list(gdir.glob("toyg.packed.*"))                      # a small scan does not pack the genome
    monkeypatch.setattr(cli, "PACKED_MIN_BP", 1)
    assert cli.main(base + ["-o", str(tmp_path / "packed")]) == 0    # packs on the device, writes the cache
    assert (gdir / "toyg.packed.chroms.tsv").exists()
    monkeypatch.setattr(cli, "PACKED_MIN_BP", 1 << 40)
    assert cli.main(base + ["-o", str(tmp_path / "cached")]) == 0    # the cache alone switches the path on
    want = _tree(tmp_path / "plain")
    assert len(want) > 10
    assert _tree(tmp_path / "packed") == want and _tree(tmp_path / "cached") == want
    n = min(_lib.device_count(), 8)
    if n >= 2:
        assert cli.main(base + ["-o", str(tmp_path / "multi"), "--gpus", str(n)]) == 0
        assert _tree(tmp_path / "multi") == want
    with pytest.raises(SystemExit):
        cli.main(base + ["-o", str(tmp_path / "toomany"), "--gpus", "64"])
    for stale in gdir.glob("toyg.packed.*"):
        stale.unlink()


def test_build_from_the_resident_genome_writes_the_same_file(workspace, tmp_path, monkeypatch):
    """`motif --build` with a large sample draws it from the packed, device-resident genome (RNG on the host in
    the reference's order, N counts and windows on the device): the built PWM file is the one the host sampler
    gives, byte for byte."""
    gdir, mdir = workspace / "toyg", workspace / "toym"
    built = mdir / "toym_toyg_pwms.motifscan"
    args = ["motif", "--build", str(mdir), "-g", str(gdir), "--n-random", "20000", "--seed", "1", "--max-n", "0"]
    for stale in gdir.glob("toyg.packed.*"):
        stale.unlink()
    assert cli.main(args) == 0
    want = built.read_text()
    monkeypatch.setattr(cli, "BUILD_RESIDENT_MIN_SAMPLES", 1)
    assert cli.main(args) == 0
    assert (gdir / "toyg.packed.chroms.tsv").exists()          # the resident path ran (and left the cache)
    assert built.read_text() == want
    assert cli.main(args + ["--n-repeat", "2"]) == 0            # also with repeats (seed, seed + 1) and the cache present
    monkeypatch.setattr(cli, "BUILD_RESIDENT_MIN_SAMPLES", 1 << 40)
    for stale in gdir.glob("toyg.packed.*"):
        stale.unlink()
    two = built.read_text()
    assert cli.main(args + ["--n-repeat", "2"]) == 0
    assert built.read_text() == two
    assert cli.main(args) == 0                                  # leave the single-repeat file for the other tests
