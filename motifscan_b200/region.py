"""GenomicRegion: 0-based half-open interval with a summit (reference region/__init__.py:17-70)."""


class GenomicRegion:
    __slots__ = ("chrom", "start", "end", "summit", "score")

    def __init__(self, chrom, start, end, summit=None, score=None):
        self.chrom = chrom
        self.start = int(start)
        self.end = int(end)
        if self.start >= self.end:
            raise ValueError(f"expect start < end, got: start={start} end={end}")
        # default summit: the midpoint (region/__init__.py:63)
        self.summit = int(summit) if summit is not None else (self.start + self.end) // 2
        self.score = score

    def __repr__(self):
        return f"GenomicRegion({self.chrom}:{self.start}-{self.end})"


def read_bed(path):
    """Minimal BED reader (chrom, start, end[, name, score]); `#`/track/browser lines skipped.
    Summit = midpoint, like the reference's 'bed' format parser."""
    out = []
    with open(path) as fh:
        for line in fh:
            if not line.strip() or line.startswith(("#", "track", "browser")):
                continue
            f = line.rstrip("\n").split("\t")
            score = None
            if len(f) > 4:
                try:
                    score = float(f[4])
                except ValueError:
                    score = None
            out.append(GenomicRegion(f[0], int(f[1]), int(f[2]), score=score))
    return out


# ---- the reference's region file formats (region/parsers.py:99-274), table-driven ---------------
def _skip_track(line):
    return line.startswith(("#", "track", "browser"))


def _skip_comment(line):
    return line.startswith("#")


def _skip_xls(line):
    return line.startswith("#") or line.split("\t")[0] == "chr"


def _opt_float(fields, k):
    try:
        return float(fields[k])
    except (TypeError, ValueError, IndexError):
        return None


def _narrowpeak_summit(f):
    rel = int(f[9])
    return None if rel == -1 else int(f[1]) + rel


# format -> (header test, line -> (chrom, start, end, summit, score)); xls tables are 1-based
REGION_FORMATS = {
    "bed": (_skip_track, lambda f: (f[0], int(f[1]), int(f[2]), None, _opt_float(f, 4))),
    "bed3-summit": (_skip_comment, lambda f: (f[0], int(f[1]), int(f[2]), int(f[3]), None)),
    "macs": (_skip_xls, lambda f: (f[0], int(f[1]) - 1, int(f[2]), int(f[4]) + int(f[1]) - 1, float(f[6]))),
    "macs2": (_skip_xls, lambda f: (f[0], int(f[1]) - 1, int(f[2]), int(f[4]) - 1, float(f[6]))),
    "narrowpeak": (_skip_track, lambda f: (f[0], int(f[1]), int(f[2]), _narrowpeak_summit(f), float(f[4]))),
    "broadpeak": (_skip_track, lambda f: (f[0], int(f[1]), int(f[2]), None, float(f[4]))),
    "manorm": (_skip_xls, lambda f: (f[0], int(f[1]) - 1, int(f[2]), int(f[3]) - 1, float(f[4]))),
}


class RegionFileFormatError(ValueError):
    pass


def load_motifscan_regions(path, format="bed"):
    """Read genomic regions in one of the reference's formats (region/__init__.py:73-94): leading
    header lines are skipped, blank lines ignored, a malformed line raises with its number."""
    try:
        is_header, parse = REGION_FORMATS[format.lower()]
    except KeyError:
        raise ValueError(f"unknown region file format: {format!r}")
    regions = []
    in_header = True
    with open(path) as fh:
        for line_num, raw in enumerate(fh, 1):
            line = raw.strip()
            if not line:
                continue
            if in_header:
                if is_header(line):
                    continue
                in_header = False
            try:
                chrom, start, end, summit, score = parse(line.split("\t"))
                regions.append(GenomicRegion(chrom, start, end, summit=summit, score=score))
            except (IndexError, ValueError, TypeError):
                raise RegionFileFormatError(f"invalid {format} line {line_num}: {line!r}")
    return regions


def generate_control_regions(n_random, regions, chrom_size, genes=None, random_seed=None):
    """`n_random` random regions per input region, same chromosome and length
    (region/utils.py:89-145, the branch without gene annotations; same `random` call sequence)."""
    import random
    if genes is not None:
        raise NotImplementedError("gene-matched control regions need the gene annotation layer (out of scope)")
    if random_seed is not None:
        random.seed(random_seed)
    controls = []
    for region in regions:
        length = region.end - region.start
        for _ in range(n_random):
            start = random.randint(0, chrom_size[region.chrom] - length)
            controls.append(GenomicRegion(region.chrom, start, start + length))
    return controls
