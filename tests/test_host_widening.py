"""CPU tests of the callers around the scan path: enrichment
statistics, region-file formats, control regions, result writers and the CLI parser, against
fixtures produced by running the reference (tests/golden/make_golden.py -> host_cases.json)."""
import json
import os
import types
from collections import namedtuple

import numpy as np
import pytest

from motifscan_b200 import io as msio
from motifscan_b200 import region as msregion
from motifscan_b200 import stats as msstats
from motifscan_b200.scanner import MotifSite, MotifSites

HERE = os.path.dirname(os.path.abspath(__file__))
Pwm = namedtuple("Pwm", ["matrix_id", "name", "length"])


@pytest.fixture(scope="module")
def host_cases():
    with open(os.path.join(HERE, "golden", "host_cases.json")) as fh:
        return json.load(fh)


def unhex(x):
    return float.fromhex(x)


def same_float(a, b):
    return (np.isnan(a) and np.isnan(b)) or np.float64(a).tobytes() == np.float64(b).tobytes()


def test_enrichment_matches_reference_bit_for_bit(host_cases):
    """Counts, fold change and both Fisher p-values equal the reference's (tolerance 0: the same
    scipy call on the same table)."""
    pwms = [Pwm(*p) for p in host_cases["enrichment"]["pwms"]]
    for case in host_cases["enrichment"]["cases"]:
        nested = lambda hits: [[[1] if h else [] for h in per] for per in hits]
        got = msstats.motif_enrichment(pwms, nested(case["hits_input"]), nested(case["hits_control"]))
        assert len(got) == len(case["results"])
        for g, w in zip(got, case["results"]):
            assert (g.name, g.n_input, g.n_control) == (w["name"], w["n_input"], w["n_control"])
            for field in ("fold_change", "p_enriched", "p_depleted", "p_corrected"):
                assert same_float(getattr(g, field), unhex(w[field])), (field, g, w)


def test_enrichment_from_array_view(host_cases):
    """The array-backed view (what Scanner.scan_motifs returns) gives the same results as the
    nested lists."""
    pwms = [Pwm(*p) for p in host_cases["enrichment"]["pwms"]]
    case = host_cases["enrichment"]["cases"][0]

    def view(hits):
        hits = np.asarray(hits)
        return types.SimpleNamespace(n_seqs=hits.shape[1], regions_with_sites=lambda: hits.sum(axis=1))

    got = msstats.motif_enrichment(pwms, view(case["hits_input"]), view(case["hits_control"]))
    for g, w in zip(got, case["results"]):
        assert same_float(g.p_enriched, unhex(w["p_enriched"])) and g.n_input == w["n_input"]


def test_region_formats(host_cases, tmp_path):
    for fmt, case in host_cases["regions"].items():
        path = tmp_path / f"regions.{fmt}"
        path.write_text(case["text"])
        got = msregion.load_motifscan_regions(str(path), fmt)
        assert [[r.chrom, r.start, r.end, r.summit, r.score] for r in got] == case["regions"], fmt
    with pytest.raises(ValueError):
        msregion.load_motifscan_regions(str(path), "nope")
    bad = tmp_path / "bad.bed"
    bad.write_text("chr1\t10\tx\n")
    with pytest.raises(msregion.RegionFileFormatError):
        msregion.load_motifscan_regions(str(bad), "bed")


def test_control_regions(host_cases):
    case = host_cases["control_regions"]
    regions = [msregion.GenomicRegion(*r) for r in case["regions"]]
    got = msregion.generate_control_regions(case["n_random"], regions, case["chrom_size"], random_seed=case["seed"])
    assert [[r.chrom, r.start, r.end] for r in got] == case["out"]


def _nested(dump):
    return [[[MotifSite(s[0], unhex(s[1]), s[2]) for s in cell] for cell in per] for per in dump]


def _as_view(nested, n_regions, lengths):
    """Nested lists -> the array-backed MotifSites the device path produces."""
    off, seq, start, score, strand = [0], [], [], [], []
    for per in nested:
        for r, cell in enumerate(per):
            for s in cell:
                seq.append(r), start.append(s.start), score.append(s.score), strand.append(1 if s.strand == "+" else 2)
        off.append(len(seq))
    res = types.SimpleNamespace(offsets=np.asarray(off, dtype=np.int64), seq_idx=np.asarray(seq, dtype=np.int32),
                                start=np.asarray(start, dtype=np.int32), score=np.asarray(score, dtype=np.float64),
                                strand=np.asarray(strand, dtype=np.int8))
    return MotifSites(res, len(nested), [0] * n_regions, lengths)


@pytest.mark.parametrize("backing", ["lists", "arrays"])
def test_writers_byte_identical(host_cases, tmp_path, backing):
    case = host_cases["writers"]
    pwms = [Pwm(*p) for p in host_cases["enrichment"]["pwms"]]
    regions = [msregion.GenomicRegion(*r) for r in case["regions"]]
    sites, ctl = _nested(case["sites"]), _nested(case["control_sites"])
    if backing == "arrays":
        sites = _as_view(sites, len(regions), [p.length for p in pwms])
        ctl = _as_view(ctl, 6, [p.length for p in pwms])
    out = str(tmp_path)
    msio.write_sites_table(out, pwms, regions, sites)
    msio.write_sites_bed(out, pwms, regions, sites)
    msio.write_enrich_table(out, msstats.motif_enrichment(pwms, sites, ctl))
    for rel, text in case["files"].items():
        with open(os.path.join(out, rel)) as fh:
            assert fh.read() == text, rel


def test_cli_parser_defaults():
    from motifscan_b200 import cli
    args = cli.make_parser().parse_args(["scan", "-i", "a.bed", "-m", "set", "-g", "hg19", "-o", "out"])
    assert (args.p_value, args.window_size, args.strand, args.n_random, args.input_format) == ("1e-4", 1000, "both", 5, "bed")
    args = cli.make_parser().parse_args(["motif", "--build", "set", "-g", "hg19"])
    assert (args.n_random, args.n_repeat, args.max_n, args.seed) == (1000000, 1, 0, None)
    with pytest.raises(SystemExit):
        cli.make_parser().parse_args(["scan", "-i", "a.bed", "-m", "set", "-g", "hg19", "-o", "out", "-p", "1e-7"])


def test_jaspar_parser_matches_the_reference_parser(tmp_path):
    """tests/golden/jaspar_cases.json: the reference's own parser (motif/__init__.py:71-140) run on well-formed
    and malformed texts -- multi-token headers (name = first token), old bare-number rows, blank lines, and
    every malformed case with the line number its error names."""
    import re
    from conftest import load_golden
    from motifscan_b200.motif import read_jaspar_pfms
    cases = load_golden("jaspar_cases.json")
    assert len(cases) >= 12
    for c in cases:
        path = tmp_path / (c["label"] + ".jaspar")
        path.write_text(c["text"])
        if "records" in c:
            got = [[mid, name, m.tolist()] for mid, name, m in read_jaspar_pfms(str(path))]
            assert got == c["records"], c["label"]
        else:
            with pytest.raises(ValueError) as err:
                read_jaspar_pfms(str(path))
            assert int(re.search(r"at line (\d+)", str(err.value)).group(1)) == c["error_line"], c["label"]
