"""CPU tests of the host-side mirror of the reference interface: motif containers, cutoff index
rule, list-based regroup/dedup helpers, pysam-free genome access, the seeded sampler."""
import os

import numpy as np
import pytest

from conftest import load_golden, unhex
from motifscan_b200.genome import Genome
from motifscan_b200.motif import MotifPwms, PositionWeightMatrix, cutoff_ranks, get_score_cutoffs, read_jaspar_pfms
from motifscan_b200.motif.matrix import pfm_to_pwm
from motifscan_b200.region import GenomicRegion
from motifscan_b200.scanner import MotifSite, deduplicate_motif_sites, make_motif_sites


@pytest.fixture(scope="module")
def toy_genome(tmp_path_factory, scanner_toy):
    d = tmp_path_factory.mktemp("genome")
    with open(d / "test.fa", "w") as fh:
        for chrom in ["chr1", "chr2", "chrX", "chrM"]:  # file order of the reference's toy genome
            fh.write(f">{chrom}\n{scanner_toy['fasta'][chrom]}\n")
    with open(d / "test_bg_freq.txt", "w") as fh:
        for b in "ACGT":
            fh.write(f"{b}\t{scanner_toy['bg_freq'][b]}\n")
    return Genome("test", str(d))


def test_get_score_cutoffs_kat():
    # reference tests/test_motif_class.py:108-122
    cut = get_score_cutoffs([list(range(0, 1000000))])
    assert len(cut) == 1 and len(cut[0]) == 5
    assert cut[0] == {'1e-2': 990000, '1e-3': 999000, '1e-4': 999900, '1e-5': 999990, '1e-6': 999999}
    with pytest.raises(ValueError):
        get_score_cutoffs([[1], [2]])
    with pytest.raises(TypeError):
        get_score_cutoffs([1, 2, 3])
    with pytest.raises(ValueError):
        get_score_cutoffs([[1, 2, 3]])


def test_cutoff_index_rule_golden():
    g = load_golden("cutoff_cases.json")
    for n, rule in g["index_rule"].items():
        assert cutoff_ranks(int(n)) == rule
    for c in g["small"]:
        cut = get_score_cutoffs([unhex(c["scores"])])[0]
        assert list(cut.keys()) == c["keys"]
        assert list(cut.values()) == unhex(c["values"])


def test_pfm_to_pwm_golden():
    for c in load_golden("matrix_cases.json"):
        pwm = PositionWeightMatrix(pfm_to_pwm(np.array(c["pfm"]), c["bg"]))
        assert pwm.matrix.tolist() == unhex(c["pwm"])
        assert pwm.max_raw_score == unhex(c["max_raw"])
        assert pwm.min_raw_score == unhex(c["min_raw"])
        seq = "ACGTNACGTNACGTNACGTNACGTNACGTN"[:pwm.length]
        assert pwm.score(seq) == unhex(c["score_first"])
    with pytest.raises(ValueError):
        pfm_to_pwm(np.array([[0, 1], [0, 1], [0, 1], [0, 1]]))


def test_pwm_score_kat():
    # reference tests/test_motif_matrix.py (score of NNN is 0, AGT 0.9186...)
    pwm = PositionWeightMatrix([[1.35, 0.21, -5.23], [0.07, -0.21, 0.6], [2.15, 2.22, -0.84], [-2.64, -1.89, 5.47]])
    assert pwm.score("NNN") == 0
    assert pwm.score("AGT") == pytest.approx(0.9186991869918698)
    with pytest.raises(ValueError):
        pwm.score("AG")


def test_motifscan_format_roundtrip(tmp_path, scanner_toy):
    pwms = MotifPwms(name="set")
    for m in scanner_toy["motifs"]:
        pwms.append(PositionWeightMatrix(unhex(m["matrix"]), name=m["name"], matrix_id=m["matrix_id"],
                                         cutoffs=unhex(m["cutoffs"])))
    path = tmp_path / "x.motifscan"
    pwms.write_motifscan_pwms(path)
    text = open(path).read()
    assert text.startswith(">MA0006.1\tAhr::Arnt\tPWM\nA [-0.85815\t-5.68647")
    assert "Cutoff_p1e-4\t0.8298548593827696\n" in text
    back = MotifPwms()
    back.read_motifscan_pwms(path)
    assert [p.matrix_id for p in back] == [m["matrix_id"] for m in scanner_toy["motifs"]]
    assert back[0].matrix.tolist() == pwms[0].matrix.tolist()
    assert back[1].cutoffs == pwms[1].cutoffs
    bad = tmp_path / "bad.motifscan"
    bad.write_text(">id\tname\tPWM\nA [1 2]\nG [1 2]\n")
    with pytest.raises(ValueError):
        MotifPwms().read_motifscan_pwms(bad)
    with pytest.raises(ValueError):
        MotifPwms(pwms=[1])


def test_read_jaspar(tmp_path):
    p = tmp_path / "m.jaspar"
    p.write_text(">MA0006.1\tAhr::Arnt\nA  [     3      0 ]\nC  [     8      0 ]\nG  [     2     23 ]\nT  [    11      1 ]\n")
    (mid, name, m), = read_jaspar_pfms(p)
    assert (mid, name) == ("MA0006.1", "Ahr::Arnt")
    assert m.tolist() == [[3, 0], [8, 0], [2, 23], [11, 1]]


def test_dedup_kat():
    # reference tests/test_scanner.py:57-73
    s = [MotifSite(1, 1, '+'), MotifSite(3, 0.8, '+'), MotifSite(1, 1, '-'), MotifSite(2, 3, '-'), MotifSite(5, 1, '+')]
    out = deduplicate_motif_sites([[s]], [3])
    assert len(out) == 1 and len(out[0]) == 1
    assert [(x.start, x.strand) for x in out[0][0]] == [(1, '+'), (2, '-'), (5, '+')]


def test_dedup_golden():
    for c in load_golden("dedup_cases.json"):
        ms = [[[MotifSite(*s) for s in seq] for seq in per] for per in c["sites"]]
        got = deduplicate_motif_sites(ms, c["lengths"])
        assert [[[list(s) for s in seq] for seq in per] for per in got] == c["dedup"]


def test_make_motif_sites():
    out = make_motif_sites([[[0, 1, 0.5, 1], [1, 0, 0.7, 2]], []], [10, 20])
    assert out == [[[MotifSite(11, 0.5, '+')], [MotifSite(20, 0.7, '-')]], [[], []]]


def test_genome_fetch_and_sizes(toy_genome):
    # reference tests/test_genome_class.py:11-23 semantics: 0-based half-open, case preserved
    assert toy_genome.chroms == ["chr1", "chr2", "chrM", "chrX"]
    assert toy_genome.chrom_sizes == {"chr1": 10, "chr2": 17, "chrM": 15, "chrX": 16}
    assert toy_genome.fetch_sequence("chr1", 2, 5) == "TtC"
    assert toy_genome.fetch_sequence("chr1", 1, 5) == "aTtC"
    assert toy_genome.fetch_sequence("chr2", 1, 150) == "AAaCCccTTtGNNNNN"  # clipped at the end
    assert toy_genome.bg_freq == {"A": 0.3, "C": 0.3, "G": 0.15, "T": 0.25}
    with pytest.raises(KeyError):
        toy_genome.fetch_sequence("chr9", 5, 123)


def test_genome_multiline_fasta(tmp_path):
    (tmp_path / "g.fa").write_text(">c1 desc\nACGTA\nCGTAC\nGT\n>c2\nNNNN\n")
    g = Genome("g", str(tmp_path), bg_freq=dict(A=.25, C=.25, G=.25, T=.25))
    assert g.chrom_sizes == {"c1": 12, "c2": 4}
    full = "ACGTACGTACGT"
    for a in range(12):
        for b in range(a, 14):
            assert g.fetch_sequence("c1", a, b) == full[a:b]
    assert os.path.exists(tmp_path / "g.fa.fai")
    g2 = Genome("g", str(tmp_path), bg_freq=g.bg_freq)  # second open goes through the .fai
    assert g2.fetch_sequence("c1", 3, 9) == full[3:9]


def test_random_sequences_golden(toy_genome):
    # reference tests/test_genome_class.py:24-28 + fixtures generated by the reference
    assert list(toy_genome.random_sequences(3, 5, random_seed=1)) == ['AAAAA', 'AaTtC', 'AAAaC']
    for c in load_golden("sampler_cases.json"):
        assert list(toy_genome.random_sequences(c["n"], c["length"], c["max_n"], c["seed"])) == c["seqs"]


def test_region():
    r = GenomicRegion("chr1", 2, 5)
    assert (r.summit, repr(r)) == (3, "GenomicRegion(chr1:2-5)")
    with pytest.raises(ValueError):
        GenomicRegion("chr1", 5, 5)


def test_scanner_init_kat(toy_genome):
    # reference tests/test_scanner.py:11-26 (sequence extraction only: no GPU needed)
    from motifscan_b200.scanner import Scanner
    regions = [GenomicRegion(chrom='chr1', start=2, end=5)]
    scanner = Scanner(genome=toy_genome, regions=regions, window_size=0)
    assert scanner.window_size == 0
    assert scanner.sequences == ['TtC']
    assert scanner.seq_starts == [2]
    assert scanner.seq_ends == [5]
    scanner = Scanner(genome=toy_genome, regions=regions, window_size=4, strand='+')
    assert scanner.sequences == ['aTtC']
    assert scanner.seq_starts == [1]
    assert scanner.seq_ends == [5]
    with pytest.raises(ValueError):
        Scanner(genome=toy_genome, regions=regions, window_size=0, strand='*')
