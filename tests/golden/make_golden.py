#!/usr/bin/env python3
"""Generate tests/golden/*.json by RUNNING THE UNMODIFIED REFERENCE in this container.

Run from the repo root (needs /root/reference and `make -C oracle`):

    python tests/golden/make_golden.py

What is executed is the reference's own code:
  * oracle/_ref/cscore*.so -- /root/reference/motifscan/motif/cscore.c compiled as-is by
    oracle/Makefile (c_score, c_scan_motif);
  * the reference's Python package imported from /root/reference (Scanner, make_motif_sites,
    deduplicate_motif_sites, get_score_cutoffs, matrix.py PFM->PPM->PWM, Genome.random_sequences),
    with (a) `motifscan.motif.cscore` pre-bound to the compiled extension above and (b) a
    minimal in-memory stand-in for `pysam.FastaFile` (pysam is not installed in this image;
    fetch = 0-based half-open slice, clipped at the chromosome end, case preserved).

The fixtures carry inputs AND outputs, so the tests that consume them never touch
/root/reference (it does not exist on the GPU box).  Floating-point outputs are stored as
`float.hex()` strings so comparisons can be bit-exact.
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("MOTIFSCAN_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

import oracle  # noqa: E402  (only for load_reference_cscore: the loader of oracle/_ref)


def hexf(x):
    return float(x).hex()


def install_pysam_stub():
    class FastaFile:
        def __init__(self, path):
            self._seqs = {}
            name = None
            with open(path) as fh:
                for line in fh:
                    line = line.rstrip("\n")
                    if line.startswith(">"):
                        name = line[1:].split()[0]
                        self._seqs[name] = []
                    elif name is not None:
                        self._seqs[name].append(line)
            self._seqs = {k: "".join(v) for k, v in self._seqs.items()}
            self.closed = False

        @property
        def references(self):
            return list(self._seqs)

        def get_reference_length(self, chrom):
            return len(self._seqs[chrom])

        def fetch(self, chrom, start, end):
            return self._seqs[chrom][start:end]

        def close(self):
            self.closed = True

    mod = types.ModuleType("pysam")
    mod.FastaFile = FastaFile
    sys.modules["pysam"] = mod


def import_reference():
    ref = oracle.load_reference_cscore()
    if ref is None:
        raise SystemExit("oracle/_ref not built: run `make -C oracle` first")
    install_pysam_stub()
    sys.path.insert(0, REFERENCE)
    # Bind the compiled extension at the import path the reference uses (scanner.py:12).
    import motifscan  # noqa: F401
    import motifscan.motif  # noqa: F401  (pure-Python part; imports genome -> pysam stub)
    sys.modules["motifscan.motif.cscore"] = ref
    motifscan.motif.cscore = ref
    return ref


def random_pwm(rng, L, kind):
    if kind == "fixed5":  # production-like: log-odds rounded to 5 dp
        p = rng.dirichlet([0.3] * 4, size=L).T
        p = np.maximum(p, 1e-3)
        p = p / p.sum(axis=0)
        bg = np.array([0.295, 0.205, 0.205, 0.295]).reshape(4, 1)
        return np.around(np.log(p / bg), 5)
    if kind == "float":  # arbitrary doubles
        return rng.normal(0, 2, size=(4, L))
    if kind == "int":  # integer PWMs like tests/test_scanner.py:35
        return rng.integers(-3, 4, size=(4, L)).astype(float)
    if kind == "neg":  # has all-negative columns (max_raw floor at 0, cscore.c:39)
        m = rng.normal(0, 2, size=(4, L))
        m[:, ::3] = -np.abs(m[:, ::3]) - 0.1
        return m
    raise ValueError(kind)


ALPHABET = np.frombuffer(b"ACGTacgtNnRYKMxz-", dtype=np.uint8)


def random_seq(rng, n, p_other):
    probs = np.array([0.2] * 4 + [0.04] * 4 + [p_other / 9] * 9)
    probs[:8] *= (1 - p_other) / probs[:8].sum()
    probs /= probs.sum()
    return bytes(ALPHABET[rng.choice(len(ALPHABET), size=n, p=probs)]).decode()


def gen_cscore_cases(ref):
    rng = np.random.default_rng(20201)
    cases = []
    # The reference's own known-answer vectors (tests/test_motif_score.py:6-32).
    kat_m = [[[1.35, 0.21, -5.23], [0.07, -0.21, 0.6], [2.15, 2.22, -0.84], [-2.64, -1.89, 5.47]]]
    for strand in (1, 2, 3):
        cases.append(dict(name=f"kat_score_s{strand}", pwms=kat_m, cutoffs=[0.2],
                          seqs=['NNN', 'AGT', 'ANT', 'CTA'], strand=strand))
        cases.append(dict(name=f"kat_scan_s{strand}", pwms=kat_m, cutoffs=[0.2],
                          seqs=['NNNAG', 'TANTCTA'], strand=strand))
    for k in range(40):
        kind = ["fixed5", "float", "int", "neg"][k % 4]
        n_pwms = int(rng.integers(1, 6))
        pwms = [random_pwm(rng, int(rng.integers(1 if k % 5 == 0 else 4, 31)), kind).tolist()
                for _ in range(n_pwms)]
        lmax = max(len(p[0]) for p in pwms)
        n_seqs = int(rng.integers(1, 7))
        p_other = [0.0, 0.02, 0.3][k % 3]
        seqs = [random_seq(rng, int(rng.integers(0 if k % 7 == 0 else lmax, 4 * lmax + 40)), p_other)
                for _ in range(n_seqs)]
        if k % 6 == 0:
            seqs.append("N" * (lmax + 5))
        if k % 9 == 0:
            seqs.append("")
        # cutoffs chosen to give sparse / dense / none / all hits
        cut_mode = k % 4
        cutoffs = []
        for p in pwms:
            if cut_mode == 0:
                cutoffs.append(float(rng.uniform(0.3, 0.8)))
            elif cut_mode == 1:
                cutoffs.append(float(rng.uniform(-0.2, 0.3)))
            elif cut_mode == 2:
                cutoffs.append(1.0)
            else:
                cutoffs.append(float(rng.uniform(-3, 0)))
        cases.append(dict(name=f"rand{k}_{kind}", pwms=pwms, cutoffs=cutoffs, seqs=seqs,
                          strand=int(rng.integers(1, 4))))
    out = []
    for c in cases:
        lmax = max(len(p[0]) for p in c["pwms"])
        sites = ref.c_scan_motif(c["pwms"], c["cutoffs"], c["seqs"], c["strand"], 2)
        c["scan"] = [[[s[0], s[1], hexf(s[2]), s[3]] for s in per] for per in sites]
        # c_score reads L columns without a length check (cscore.c:195): only call it on
        # sequences long enough for every motif.
        long_seqs = [s for s in c["seqs"] if len(s) >= lmax]
        c["score_seqs"] = long_seqs
        if long_seqs:
            sc = ref.c_score(c["pwms"], long_seqs, c["strand"], 2)
            c["score"] = [[hexf(v) for v in row] for row in sc]
        else:
            c["score"] = [[] for _ in c["pwms"]]
        c["pwms"] = [[[hexf(v) for v in row] for row in p] for p in c["pwms"]]
        c["cutoffs"] = [hexf(v) for v in c["cutoffs"]]
        out.append(c)
    return out


def gen_scanner_cases():
    """Toy genome + the shipped built PWM file through the reference's Scanner."""
    from motifscan.genome import Genome
    from motifscan.motif import MotifPwms
    from motifscan.motif.matrix import PositionWeightMatrix
    from motifscan.region import GenomicRegion, load_motifscan_regions
    from motifscan.scanner import Scanner

    data = os.path.join(REFERENCE, "tests", "data")
    genome = Genome(name="test", path=os.path.join(data, "genomes", "test"))
    fasta = {c: genome.fetch_sequence(c, 0, genome.chrom_sizes[c]) for c in genome.chroms}
    pwms = MotifPwms()
    pwms.read_motifscan_pwms(os.path.join(data, "motifs", "test", "test_pwms.motifscan"))
    regions = load_motifscan_regions(os.path.join(data, "regions", "test_regions.bed"), "bed")
    regions = [r for r in regions if r.chrom in genome.chrom_sizes]  # drop the chr9 line
    extra = [GenomicRegion("chrX", 0, 16), GenomicRegion("chrM", 2, 14), GenomicRegion("chr2", 3, 12)]
    regions = regions + extra
    motif_dump = [dict(matrix_id=p.matrix_id, name=p.name, matrix=[[hexf(v) for v in row] for row in p.matrix.tolist()],
                       cutoffs={k: hexf(v) for k, v in p.cutoffs.items()}) for p in pwms]
    cases = []
    for window in (1000, 0, 8):
        for p_value in ("1e-2", "1e-3", "1e-4"):
            for strand in ("both", "+", "-"):
                for remove_dup in (True, False):
                    sc = Scanner(genome=genome, regions=regions, window_size=window, strand=strand,
                                 p_value=p_value, remove_dup=remove_dup, n_threads=1)
                    res = sc.scan_motifs(pwms)
                    cases.append(dict(
                        window_size=window, p_value=p_value, strand=strand, remove_dup=remove_dup,
                        sequences=sc.sequences, seq_starts=sc.seq_starts, seq_ends=sc.seq_ends,
                        sites=[[[[s.start, hexf(s.score), s.strand] for s in per_seq]
                                for per_seq in per_pwm] for per_pwm in res]))
    # the reference's own integer-PWM scanner test (tests/test_scanner.py:29-54)
    ipwms = MotifPwms()
    ipwms.append(PositionWeightMatrix([[1, 0], [0, 1], [0, 0], [1, 0]], cutoffs={'1e-3': 0.5, '1e-4': 1}))
    ireg = [GenomicRegion(chrom='chr1', start=2, end=5)]
    int_cases = []
    for p_value, remove_dup in (("1e-4", True), ("1e-3", True), ("1e-3", False)):
        sc = Scanner(genome=genome, regions=ireg, window_size=4, p_value=p_value, remove_dup=remove_dup)
        res = sc.scan_motifs(ipwms)
        int_cases.append(dict(p_value=p_value, remove_dup=remove_dup, sequences=sc.sequences,
                              seq_starts=sc.seq_starts,
                              sites=[[[[s.start, hexf(s.score), s.strand] for s in per_seq]
                                      for per_seq in per_pwm] for per_pwm in res]))
    return dict(fasta=fasta, chroms=genome.chroms, bg_freq=genome.bg_freq,
                regions=[[r.chrom, r.start, r.end, r.summit] for r in regions],
                motifs=motif_dump, cases=cases, int_cases=int_cases)


def gen_dedup_cases():
    from motifscan.scanner import MotifSite, deduplicate_motif_sites
    rng = np.random.default_rng(77)
    cases = []
    for k in range(30):
        n_pwms = int(rng.integers(1, 4))
        lengths = [int(rng.integers(1, 12)) for _ in range(n_pwms)]
        ms = []
        for m in range(n_pwms):
            per_seq = []
            for _ in range(int(rng.integers(1, 4))):
                n = int(rng.integers(0, 14))
                starts = np.sort(rng.integers(0, 30, size=n))
                sites = []
                for s in starts:
                    # quantised scores make ties (the `>=` rule, scanner.py:163) likely
                    sites.append(MotifSite(int(s), float(rng.integers(0, 4)) / 4, '+'))
                    if rng.random() < 0.5:
                        sites.append(MotifSite(int(s), float(rng.integers(0, 4)) / 4, '-'))
                # reference order within a sequence: start asc, '+' before '-' at equal start;
                # repeated starts on one strand cannot occur in real output -> drop them
                seen = set()
                uniq = []
                for s in sites:
                    if (s.start, s.strand) not in seen:
                        seen.add((s.start, s.strand))
                        uniq.append(s)
                per_seq.append(uniq)
            ms.append(per_seq)
        res = deduplicate_motif_sites(ms, lengths)
        cases.append(dict(lengths=lengths,
                          sites=[[[list(s) for s in seq] for seq in per] for per in ms],
                          dedup=[[[list(s) for s in seq] for seq in per] for per in res]))
    return cases


def gen_cutoff_cases():
    from motifscan.motif import get_score_cutoffs
    rng = np.random.default_rng(5)
    # small explicit cases keep the score vectors in the fixture
    small = []
    for n in (100, 150, 1000, 12345):
        scores = np.round(rng.normal(0, 0.3, size=n), 2).tolist()
        cut = get_score_cutoffs([list(scores)])[0]
        small.append(dict(scores=[hexf(v) for v in scores], keys=list(cut.keys()),
                          values=[hexf(v) for v in cut.values()]))
    # index rule as a table: n -> {key: index}
    index_rule = {}
    for n in (100, 101, 999, 1000, 4321, 10000, 99999, 100000, 250000, 999999, 1000000,
              1200000, 9999999, 10000000, 12345678):
        cut = get_score_cutoffs([list(range(n))])[0]  # sorted desc -> value v sits at index n-1-v
        index_rule[str(n)] = {k: int(n - 1 - v) for k, v in cut.items()}
    return dict(small=small, index_rule=index_rule)


def gen_matrix_cases():
    from motifscan.motif import MotifPfms
    data = os.path.join(REFERENCE, "tests", "data")
    pfms = MotifPfms()
    pfms.read_pfms(os.path.join(data, "motifs", "test", "test_pfms.jaspar"), format="jaspar")
    bgs = [None, dict(A=0.3, C=0.3, G=0.15, T=0.25), dict(A=0.295, C=0.205, G=0.205, T=0.295)]
    out = []
    for pfm in pfms:
        for bg in bgs:
            pwm = pfm.to_ppm().to_pwm(bg)
            out.append(dict(matrix_id=pfm.matrix_id, name=pfm.name, pfm=pfm.matrix.tolist(), bg=bg,
                            pwm=[[hexf(v) for v in row] for row in pwm.matrix.tolist()],
                            max_raw=hexf(pwm.max_raw_score), min_raw=hexf(pwm.min_raw_score),
                            score_first=hexf(pwm.score("ACGTNACGTNACGTNACGTNACGTNACGTN"[:pwm.length]))))
    return out


def gen_sampler_cases():
    from motifscan.genome import Genome
    data = os.path.join(REFERENCE, "tests", "data")
    genome = Genome(name="test", path=os.path.join(data, "genomes", "test"))
    out = []
    for n, length, max_n, seed in ((3, 5, 0, 1), (10, 4, 0, 7), (10, 6, 2, 3), (25, 3, 1, 11)):
        out.append(dict(n=n, length=length, max_n=max_n, seed=seed,
                        seqs=list(genome.random_sequences(n, length, max_n, seed))))
    return out


def gen_build_case(ref):
    """`motif --build` arithmetic (cli/motif.py:119-153) on the toy genome: PFM -> PWM with the
    toy bg, sample, c_score(..., 3), get_score_cutoffs, mean + round 8."""
    from motifscan.genome import Genome
    from motifscan.motif import MotifPfms, get_score_cutoffs
    data = os.path.join(REFERENCE, "tests", "data")
    genome = Genome(name="test", path=os.path.join(data, "genomes", "test"))
    pfms = MotifPfms()
    pfms.read_pfms(os.path.join(data, "motifs", "test", "test_pfms.jaspar"), format="jaspar")
    pwms = [pfm.to_ppm().to_pwm(genome.bg_freq) for pfm in pfms]
    # the toy chromosomes are shorter than MA0854.1 (17) + anything; keep only the 6-mer
    pwms = [p for p in pwms if p.length <= 6]
    max_length = max(p.length for p in pwms)
    n_random, n_repeat, seed0, max_n = 500, 2, 4, 6
    cutoffs_all = []
    for i in range(n_repeat):
        seqs = list(genome.random_sequences(n_random, max_length, max_n, seed0 + i))
        scores = ref.c_score([p.matrix.tolist() for p in pwms], seqs, 3, 1)
        cutoffs_all.append(get_score_cutoffs(scores))
    final = []
    for i in range(len(pwms)):
        acc = {}
        for rep in cutoffs_all:
            for k, v in rep[i].items():
                acc.setdefault(k, []).append(v)
        final.append({k: hexf(np.around(np.mean(v), 8)) for k, v in acc.items()})
    return dict(n_random=n_random, n_repeat=n_repeat, seed=seed0, max_n=max_n, max_length=max_length,
                pwms=[[[hexf(v) for v in row] for row in p.matrix.tolist()] for p in pwms],
                cutoffs=final)


def gen_host_cases():
    """stats.motif_enrichment, the region-file parsers, the control-region generator and the result
    writers of the reference, run on small inputs (inputs and outputs both stored)."""
    import random
    import tempfile
    from collections import namedtuple
    from motifscan.io import write_enrich_table, write_sites_bed, write_sites_table
    from motifscan.region import GenomicRegion, load_motifscan_regions
    from motifscan.region.utils import generate_control_regions
    from motifscan.scanner import MotifSite
    from motifscan.stats import motif_enrichment

    rng = np.random.default_rng(77)
    Pwm = namedtuple("Pwm", ["matrix_id", "name", "length"])
    pwms = [Pwm(f"MA{k:04d}.1", f"TF{k}" if k % 3 else f"TF{k}::X-{k}", int(rng.integers(6, 20))) for k in range(9)]

    def random_sites(n_regions, density):
        out = []
        for _ in pwms:
            per = []
            for r in range(n_regions):
                n = int(rng.poisson(density))
                cell = [MotifSite(int(rng.integers(0, 5000)), float(rng.uniform(0.7, 1.0)), "+-"[int(rng.integers(2))])
                        for _ in range(n)]
                per.append(sorted(cell, key=lambda s: s.start))
            out.append(per)
        return out

    cases = {}
    # --- enrichment -----------------------------------------------------------------------------
    enrich = []
    for n_in, n_ctl, d_in, d_ctl in [(40, 60, 0.5, 0.3), (25, 25, 0.05, 0.6), (30, 10, 1.5, 0.0), (8, 200, 0.2, 0.2)]:
        sites, ctl = random_sites(n_in, d_in), random_sites(n_ctl, d_ctl)
        res = motif_enrichment(pwms, sites, ctl)
        enrich.append(dict(
            n_input_total=n_in, n_control_total=n_ctl,
            hits_input=[[int(len(c) > 0) for c in per] for per in sites],
            hits_control=[[int(len(c) > 0) for c in per] for per in ctl],
            results=[dict(name=r.name, n_input=int(r.n_input), n_control=int(r.n_control),
                          fold_change=hexf(r.fold_change), p_enriched=hexf(r.p_enriched),
                          p_depleted=hexf(r.p_depleted), p_corrected=hexf(r.p_corrected)) for r in res]))
    cases["enrichment"] = dict(pwms=[list(p) for p in pwms], cases=enrich)
    # --- region parsers -------------------------------------------------------------------------
    region_dir = os.path.join(REFERENCE, "tests", "data", "regions")
    files = {"bed": "test_regions.bed", "bed3-summit": "test_regions_summit.bed", "macs": "test_regions_macs.xls",
             "macs2": "test_regions_macs2.xls", "narrowpeak": "test_regions.narrowPeak",
             "broadpeak": "test_regions.broadPeak", "manorm": "test_regions_manorm.xls"}
    parsed = {}
    for fmt, fname in files.items():
        path = os.path.join(region_dir, fname)
        regions = load_motifscan_regions(path, fmt)
        parsed[fmt] = dict(text=open(path).read(),
                           regions=[[r.chrom, r.start, r.end, r.summit, r.score] for r in regions])
    cases["regions"] = parsed
    # --- control regions ------------------------------------------------------------------------
    base = [GenomicRegion("chr1", 100, 350), GenomicRegion("chr2", 5, 1005), GenomicRegion("chrX", 40, 41)]
    sizes = {"chr1": 5000, "chr2": 1200, "chrX": 300}
    ctl = generate_control_regions(3, base, sizes, genes=None, random_seed=11)
    cases["control_regions"] = dict(regions=[[r.chrom, r.start, r.end] for r in base], chrom_size=sizes, n_random=3,
                                    seed=11, out=[[r.chrom, r.start, r.end] for r in ctl])
    # --- writers --------------------------------------------------------------------------------
    regions = [GenomicRegion("chr1", 10, 510), GenomicRegion("chr1", 700, 1300), GenomicRegion("chr7", 0, 90),
               GenomicRegion("chrX", 5, 45)]
    sites, ctl_sites = random_sites(len(regions), 0.8), random_sites(6, 0.4)
    with tempfile.TemporaryDirectory() as tmp:
        write_sites_table(tmp, pwms, regions, sites)
        write_sites_bed(tmp, pwms, regions, sites)
        write_enrich_table(tmp, motif_enrichment(pwms, sites, ctl_sites))
        files_out = {}
        for root, _, names in os.walk(tmp):
            for n in names:
                full = os.path.join(root, n)
                files_out[os.path.relpath(full, tmp)] = open(full).read()
    dump = lambda nested: [[[[s.start, hexf(s.score), s.strand] for s in cell] for cell in per] for per in nested]
    cases["writers"] = dict(regions=[[r.chrom, r.start, r.end] for r in regions], sites=dump(sites),
                            control_sites=dump(ctl_sites), files=files_out)
    return cases


JASPAR_TEXTS = {
    "new_format": ">MA0006.1\tAhr::Arnt\nA  [     3      0      0 ]\nC  [     8      0     23 ]\nG  [     2     23      0 ]\nT  [    11      1      1 ]\n",
    "two_records_blank_lines": "\n>M1 first\nA [1 2]\nC [3 4]\n\nG [5 6]\nT [7 8]\n\n>M2\tsecond\nA [9]\nC [8]\nG [7]\nT [6]\n\n",
    "multi_token_header": ">MA0001.1 AGL3 extra tokens here\nA [0 3]\nC [94 75]\nG [1 0]\nT [2 19]\n",
    "header_without_name": ">MA0002.1\nA [1]\nC [2]\nG [3]\nT [4]\n",
    "header_with_space_after_gt": ">  MA0003.1   NAME\nA [1]\nC [2]\nG [3]\nT [4]\n",
    "old_format_rows": ">MA0004.1 Arnt\n4 19 0 0\n16 0 20 0\n0 1 0 20\n0 0 0 0\n",
    "mixed_old_and_new_rows": ">MA0005.1 x\nA [1 2]\n3 4\nG [5 6]\n7 8\n",
    "rows_out_of_order": ">M x\nC [1]\nA [2]\nG [3]\nT [4]\n",
    "row_before_header": "A [1]\n>M x\nA [1]\nC [2]\nG [3]\nT [4]\n",
    "header_inside_matrix": ">M x\nA [1]\nC [2]\n>N y\nG [3]\nT [4]\n",
    "non_integer_count": ">M x\nA [1.5]\nC [2]\nG [3]\nT [4]\n",
    "truncated_last_record": ">M x\nA [1]\nC [2]\nG [3]\nT [4]\n>N y\nA [1]\nC [2]\n",
    "empty_file": "",
    "fifth_row": ">M x\nA [1]\nC [2]\nG [3]\nT [4]\nA [5]\n",
}


def gen_jaspar_cases():
    """The reference's own JASPAR PFM parser (motif/__init__.py:71-140) on well-formed and malformed texts:
    parsed (id, name, values) records, or the line number its PfmsJasparFormatError names."""
    import tempfile
    from motifscan.exceptions import PfmsJasparFormatError
    from motifscan.motif import MotifPfms
    cases = []
    for label, text in JASPAR_TEXTS.items():
        with tempfile.NamedTemporaryFile("w", suffix=".jaspar", delete=False) as fh:
            fh.write(text)
            path = fh.name
        try:
            pfms = MotifPfms._parse_jaspar_pfms(path)
            cases.append(dict(label=label, text=text,
                              records=[[p.matrix_id, p.name, np.asarray(p.matrix).tolist()] for p in pfms]))
        except PfmsJasparFormatError as e:
            cases.append(dict(label=label, text=text, error_line=int(str(e).split(' at line ')[1].split(':')[0])))
        except ValueError as e:     # the matrix class rejects e.g. ragged rows after parsing
            cases.append(dict(label=label, text=text, error=type(e).__name__))
        finally:
            os.unlink(path)
    return cases


def main():
    ref = import_reference()
    if "--jaspar-only" in sys.argv:
        path = os.path.join(HERE, "jaspar_cases.json")
        with open(path, "w") as fh:
            json.dump(gen_jaspar_cases(), fh, separators=(",", ":"))
        print(f"wrote {path} ({os.path.getsize(path)} bytes)")
        return
    out = {
        "jaspar_cases.json": gen_jaspar_cases(),
        "cscore_cases.json": gen_cscore_cases(ref),
        "scanner_toy.json": gen_scanner_cases(),
        "dedup_cases.json": gen_dedup_cases(),
        "cutoff_cases.json": gen_cutoff_cases(),
        "matrix_cases.json": gen_matrix_cases(),
        "sampler_cases.json": gen_sampler_cases(),
        "build_case.json": gen_build_case(ref),
        "host_cases.json": gen_host_cases(),
    }
    for name, obj in out.items():
        path = os.path.join(HERE, name)
        with open(path, "w") as fh:
            json.dump(obj, fh, separators=(",", ":"))
        print(f"wrote {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    main()
