"""Pin the oracle (CPU restatement) against the reference: its own known-answer tests, the
golden fixtures produced by running the unmodified reference (tests/golden/make_golden.py), and
-- when oracle/_ref was built in this checkout -- live differential runs against it."""
import numpy as np
import pytest

import oracle
from conftest import load_golden, unhex

KAT_M = [[[1.35, 0.21, -5.23], [0.07, -0.21, 0.6], [2.15, 2.22, -0.84], [-2.64, -1.89, 5.47]]]


def test_kat_c_score():
    # reference tests/test_motif_score.py:6-20
    seqs = ['NNN', 'AGT', 'ANT', 'CTA']
    assert oracle.c_score(KAT_M, seqs, 1)[0] == pytest.approx(
        [0.0, 0.9186991869918698, 0.693089430894309, -0.7164634146341464])
    assert oracle.c_score(KAT_M, seqs, 2)[0] == pytest.approx(
        [0.0, 0.6717479674796748, 0.693089430894309, -0.3323170731707317])
    assert oracle.c_score(KAT_M, seqs, 3)[0] == pytest.approx(
        [0.0, 0.9186991869918698, 0.693089430894309, -0.3323170731707317])


def test_kat_c_scan_motif():
    # reference tests/test_motif_score.py:23-32
    sites = oracle.c_scan_motif(KAT_M, [0.2], ['NNNAG', 'TANTCTA'], 3)
    assert len(sites) == 1 and len(sites[0]) == 4
    assert sites[0][0] == pytest.approx([1, 1, 0.693089430894309, 1])
    assert sites[0][1] == pytest.approx([1, 1, 0.693089430894309, 2])
    assert sites[0][2] == pytest.approx([1, 2, 0.23983739837398374, 2])
    assert sites[0][3] == pytest.approx([1, 3, 0.266260162601626, 1])


def test_encode_table():
    # cscore.c:81-114: only ACGT/acgt map to codes; everything else is -1
    s = bytes(range(1, 256)).decode("latin-1").encode("utf-8")
    codes = oracle.encode(s)
    raw = np.frombuffer(s, dtype=np.uint8)
    expect = np.full(len(raw), -1, dtype=np.int8)
    for ch, v in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
        expect[raw == ch] = v
    assert np.array_equal(codes, expect)


def test_golden_cscore_bit_exact(cscore_cases):
    for c in cscore_cases:
        got = oracle.c_scan_motif(c["pwms"], c["cutoffs"], c["seqs"], c["strand"])
        assert got == c["scan"], c["name"]  # ints and doubles compared exactly
        if c["score_seqs"]:
            sc = oracle.c_score(c["pwms"], c["score_seqs"], c["strand"])
            assert sc == c["score"], c["name"]


def test_score_short_sequence_is_an_error():
    with pytest.raises(ValueError):
        oracle.c_score(KAT_M, ["AC"], 3)


def test_max_raw_floor_at_zero():
    # cscore.c:39: col_max starts at 0 -> an all-negative column contributes 0
    assert oracle.max_raw_score([[-1, 2], [-2, 1], [-3, 0], [-4, -1]]) == 2.0


def test_golden_dedup():
    for c in load_golden("dedup_cases.json"):
        ms = [[[oracle.MotifSite(*s) for s in seq] for seq in per] for per in c["sites"]]
        got = oracle.deduplicate_motif_sites(ms, c["lengths"])
        assert [[[list(s) for s in seq] for seq in per] for per in got] == c["dedup"]


def test_kat_dedup():
    # reference tests/test_scanner.py:57-73
    S = oracle.MotifSite
    out = oracle.deduplicate_motif_sites(
        [[[S(1, 1, '+'), S(3, 0.8, '+'), S(1, 1, '-'), S(2, 3, '-'), S(5, 1, '+')]]], [3])
    assert [(s.start, s.strand) for s in out[0][0]] == [(1, '+'), (2, '-'), (5, '+')]


def test_kat_score_cutoffs():
    # reference tests/test_motif_class.py:108-122
    cut = oracle.get_score_cutoffs([list(range(1000000))])[0]
    assert cut == {'1e-2': 990000, '1e-3': 999000, '1e-4': 999900, '1e-5': 999990, '1e-6': 999999}
    with pytest.raises(ValueError):
        oracle.get_score_cutoffs([list(range(99))])


def test_golden_score_cutoffs():
    g = load_golden("cutoff_cases.json")
    for c in g["small"]:
        cut = oracle.get_score_cutoffs([unhex(c["scores"])])[0]
        assert list(cut.keys()) == c["keys"]
        assert list(cut.values()) == unhex(c["values"])
    for n, rule in g["index_rule"].items():
        n = int(n)
        n_bits = min(len(str(n)), 7)
        assert {f"1e-{e}": int(n * 0.1 ** e) - 1 for e in range(2, n_bits)} == rule


def test_golden_pfm_to_pwm():
    for c in load_golden("matrix_cases.json"):
        pwm = oracle.pfm_to_pwm(c["pfm"], c["bg"])
        assert pwm.tolist() == unhex(c["pwm"])
        seq = "ACGTNACGTNACGTNACGTNACGTNACGTN"[:pwm.shape[1]]
        assert oracle.pwm_score(pwm, seq) == unhex(c["score_first"])


def test_golden_scanner_regroup_and_dedup(scanner_toy):
    """Toy genome through oracle scan + regroup + dedup == the reference Scanner's output."""
    motifs = scanner_toy["motifs"]
    pwms = [unhex(m["matrix"]) for m in motifs]
    lengths = [len(p[0]) for p in pwms]
    strand_arg = {"+": 1, "-": 2, "both": 3}
    for c in scanner_toy["cases"]:
        cutoffs = [float.fromhex(m["cutoffs"][c["p_value"]]) for m in motifs]
        sites = oracle.c_scan_motif(pwms, cutoffs, c["sequences"], strand_arg[c["strand"]])
        ms = oracle.make_motif_sites(sites, c["seq_starts"])
        if c["remove_dup"]:
            ms = oracle.deduplicate_motif_sites(ms, lengths)
        got = [[[[s.start, s.score.hex(), s.strand] for s in seq] for seq in per] for per in ms]
        assert got == c["sites"]


def test_live_differential_vs_reference_extension():
    """When the unmodified reference extension is present (oracle/_ref), compare on fresh
    random inputs -- bit-exact, including order."""
    ref = oracle.load_reference_cscore()
    if ref is None:
        pytest.skip("oracle/_ref not built in this checkout")
    rng = np.random.default_rng(99)
    alphabet = np.frombuffer(b"ACGTacgtNn", dtype=np.uint8)
    for k in range(25):
        pwms = [np.around(rng.normal(0, 2, size=(4, int(rng.integers(3, 31)))), 5).tolist()
                for _ in range(int(rng.integers(1, 5)))]
        lmax = max(len(p[0]) for p in pwms)
        seqs = [bytes(alphabet[rng.integers(0, 10, size=int(rng.integers(lmax, 300)))]).decode()
                for _ in range(int(rng.integers(1, 6)))]
        cutoffs = [float(rng.uniform(0.0, 0.7)) for _ in pwms]
        strand = int(rng.integers(1, 4))
        assert oracle.c_scan_motif(pwms, cutoffs, seqs, strand, 2) == \
            ref.c_scan_motif(pwms, cutoffs, seqs, strand, 2)
        assert oracle.c_score(pwms, seqs, strand) == ref.c_score(pwms, seqs, strand, 2)
