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
