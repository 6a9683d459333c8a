"""motifscan_b200.motif -- the motif-side host objects of the scan path.

Mirrors the names the reference's callers use (motifscan/motif/__init__.py, matrix.py):
`PositionWeightMatrix` (.matrix, .length, .cutoffs, .max_raw_score, .score), `MotifPwms`
(a list of PWMs with the `.motifscan` text format), `get_score_cutoffs`.  The extension-module
mirror lives in `motifscan_b200.motif.cscore`.
"""
import re

import numpy as np

from .matrix import BASES, pfm_to_pwm  # noqa: F401

__all__ = ["PositionWeightMatrix", "MotifPwms", "get_score_cutoffs", "cutoff_ranks",
           "read_jaspar_pfms"]


class PositionWeightMatrix:
    """4 x L log-odds matrix, rows A, C, G, T (reference matrix.py:174-240)."""

    def __init__(self, values, name=None, matrix_id=None, cutoffs=None):
        m = np.asarray(values)
        if m.ndim != 2 or m.shape[0] != 4 or m.shape[1] == 0:
            raise ValueError("a position matrix needs 4 rows (A, C, G, T) and at least one column")
        if not np.issubdtype(m.dtype, np.number):
            raise ValueError("position matrix values must be numeric")
        self.matrix = m
        self.name = name
        self.matrix_id = matrix_id
        self.cutoffs = cutoffs

    @property
    def length(self):
        return self.matrix.shape[1]

    def __len__(self):
        return self.length

    def set_cutoff(self, p_value, cutoff):
        if self.cutoffs is None:
            self.cutoffs = {}
        self.cutoffs[p_value] = cutoff

    @property
    def max_raw_score(self):
        # matrix.py:202-207 (numpy column maxima; NOT floored at 0 like cscore.c:36-48)
        return self.matrix.max(axis=0).sum()

    @property
    def min_raw_score(self):
        return self.matrix.min(axis=0).sum()

    def score(self, sequence):
        """Forward-strand score of one window, N skipped (matrix.py:216-240)."""
        if len(sequence) != self.length:
            raise ValueError("sequence should have the same length as the PWM")
        raw = 0
        for col, nt in enumerate(sequence.upper()):
            row = BASES.find(nt)
            if row >= 0:
                raw += self.matrix[row, col]
        return raw / self.max_raw_score


_HEADER = re.compile(r"^>(\S+)\t(\S+)\tPWM$")
_ROW = re.compile(r"^([ACGT]) \[(.+)\]$")
_CUTOFF = re.compile(r"^Cutoff_p(\S+)\t(\S+)")


class MotifPwms(list):
    """A set of PWMs; reads/writes the reference's `.motifscan` text format
    (motif/__init__.py:192-319): `>id<TAB>name<TAB>PWM`, four `B [v v ...]` rows with `%8.5f`
    values, then `Cutoff_p{p}<TAB>{cutoff}` lines."""

    def __init__(self, pwms=None, name=None, genome=None):
        super().__init__()
        self.name = name
        self.genome = genome
        for pwm in (pwms or []):
            if not isinstance(pwm, PositionWeightMatrix):
                raise ValueError(f"invalid PWM item: {pwm!r}")
            self.append(pwm)

    def write_motifscan_pwms(self, path):
        with open(path, "w") as out:
            for pwm in self:
                out.write(f">{pwm.matrix_id}\t{pwm.name}\tPWM\n")
                for base, row in zip(BASES, pwm.matrix):
                    out.write(base + " [" + "\t".join(f"{v:8.5f}" for v in row) + "]\n")
                for p, cutoff in (pwm.cutoffs or {}).items():
                    out.write(f"Cutoff_p{p}\t{cutoff}\n")

    def read_motifscan_pwms(self, path):
        """Strict reader: a record is 1 header, rows A C G T in order, >= 1 cutoff line."""
        def fail(num, line):
            raise ValueError(f"invalid MotifScan PWMs format at line {num}: {line!r}")

        records = []
        cur = None
        num = 0
        with open(path) as fh:
            for num, raw in enumerate(fh, 1):
                line = raw.strip()
                if not line:
                    continue
                h, r, c = _HEADER.match(line), _ROW.match(line), _CUTOFF.match(line)
                if h:
                    if cur is not None and (len(cur["rows"]) != 4 or not cur["cutoffs"]):
                        fail(num, line)
                    cur = dict(id=h.group(1), name=h.group(2), rows=[], cutoffs={})
                    records.append(cur)
                elif r:
                    if cur is None or len(cur["rows"]) >= 4 or cur["cutoffs"] or \
                            r.group(1) != BASES[len(cur["rows"])]:
                        fail(num, line)
                    try:
                        cur["rows"].append([float(v) for v in r.group(2).split()])
                    except ValueError:
                        fail(num, line)
                elif c:
                    if cur is None or len(cur["rows"]) != 4:
                        fail(num, line)
                    cur["cutoffs"][c.group(1)] = float(c.group(2))
                else:
                    fail(num, line)
        if cur is not None and (len(cur["rows"]) != 4 or not cur["cutoffs"]):
            fail(num + 1, "")
        for rec in records:
            self.append(PositionWeightMatrix(rec["rows"], name=rec["name"], matrix_id=rec["id"],
                                             cutoffs=rec["cutoffs"]))


_PFM_HEADER = re.compile(r"^>\s*(\S+)(\s+(\S+))?")          # id, optional name = the next token only
_PFM_ROW_NEW = re.compile(r"\s*([ACGT])\s*\[\s*(.+)\s*\]")     # `A  [ 3 0 0 ]`
_PFM_ROW_OLD = re.compile(r"\s*(.+)\s*")                        # bare counts, rows in A C G T order


def read_jaspar_pfms(path):
    """JASPAR text PFMs with the reference's grammar (motif/__init__.py:71-140): a header `>ID [NAME]`
    (the name is the first token after the id, so it can be written back as `>ID<TAB>NAME<TAB>PWM`), then
    exactly four count rows, either `A [ n n ... ]` (rows must come as A, C, G, T) or the old bare-number
    form; blank lines anywhere.  A row where a header is due, a header inside a matrix, a non-integer count
    or a truncated last record raise ValueError naming the line (the reference raises PfmsJasparFormatError
    at the same line).  Returns a list of (matrix_id, name, int array 4 x L)."""
    out = []
    want_header = True
    num = 0
    matrix_id = name = None
    rows = []
    with open(path) as fh:
        for num, raw in enumerate(fh, 1):
            line = raw.strip()
            if not line:
                continue
            head = _PFM_HEADER.match(line)
            if bool(head) != want_header:
                raise ValueError(f"invalid JASPAR PFM record at line {num}: {line!r}")
            if head:
                matrix_id, name, rows = head.group(1), head.group(3), []
                want_header = False
                continue
            new = _PFM_ROW_NEW.match(line)
            if new:
                if new.group(1) != BASES[len(rows)]:
                    raise ValueError(f"invalid JASPAR PFM row at line {num}: {line!r}")
                tokens = new.group(2).split()
            else:
                tokens = _PFM_ROW_OLD.match(line).group(1).split()
            try:
                rows.append([int(v) for v in tokens])
            except ValueError:
                raise ValueError(f"invalid JASPAR PFM row at line {num}: {line!r}") from None
            if len(rows) == 4:
                out.append((matrix_id, name, np.asarray(rows, dtype=np.int64)))
                want_header = True
    if not want_header:
        raise ValueError(f"invalid JASPAR PFM record at line {num + 1}: the last matrix is incomplete")
    return out


def cutoff_ranks(n_scores):
    """Indices into the descending-sorted scores the reference reads (motif/__init__.py:394-399):
    {'1e-e': int(n * 0.1 ** e) - 1 for e = 2 .. min(len(str(n)), 7) - 1}, evaluated with Python
    float arithmetic exactly as the reference does."""
    if n_scores < 100:
        raise ValueError("each motif must have at least 100 sampling scores")
    n_bits = min(len(str(n_scores)), 7)
    return {f"1e-{e}": int(n_scores * 0.1 ** e) - 1 for e in range(2, n_bits)}


def get_score_cutoffs(sampling_scores):
    """Score cutoffs from background score distributions (motif/__init__.py:378-401).

    `sampling_scores` is (n_motifs, n_sampling) array-like (lists of lists as the reference's
    c_score returns, or a 2-D array).  Returns a list of {p_value: cutoff} dicts.  Uses a partial
    sort instead of the reference's full `list.sort`; the selected order statistics are the same
    values."""
    out = []
    for scores in sampling_scores:
        n = len(scores)  # TypeError for non-nested input, like the reference
        ranks = cutoff_ranks(n)
        a = np.asarray(scores)
        idx = sorted(set(ranks.values()))
        # k-th largest == element n-1-k of the ascending order
        part = np.partition(a, [n - 1 - k for k in idx])
        out.append({p: part[n - 1 - k].item() for p, k in ranks.items()})
    return out
