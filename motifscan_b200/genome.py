"""pysam-free genome access for the scan path (reference motifscan/genome/__init__.py:61-176).

`Genome` reads a FASTA file through its `.fai` index (built on first use when missing) by
memory-mapping the file, so `fetch_sequence` is a slice + newline strip.  Semantics follow
`pysam.FastaFile.fetch` as the reference uses it: 0-based half-open coordinates, case preserved,
`end` clipped at the chromosome length.
"""
import mmap
import os

import numpy as np

BASES = "ACGT"


class Genome:
    def __init__(self, name, path=None, fasta=None, bg_freq=None):
        self.name = name
        self.path = path
        self._fasta_path = fasta or os.path.join(path, f"{name}.fa")
        if not os.path.isfile(self._fasta_path):
            raise FileNotFoundError(f"genome sequence file not found: {self._fasta_path}")
        self._index = self._load_index()
        self._fh = open(self._fasta_path, "rb")
        self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ) \
            if os.path.getsize(self._fasta_path) else b""
        if bg_freq is not None:
            self.bg_freq = dict(bg_freq)
        else:
            bg_path = os.path.join(path, f"{name}_bg_freq.txt") if path else None
            self.bg_freq = read_bg_freq(bg_path) if bg_path and os.path.isfile(bg_path) else None
        self._chroms = None
        self._chrom_sizes = None

    # -- index ---------------------------------------------------------------------------------
    def _load_index(self):
        fai = self._fasta_path + ".fai"
        if not os.path.isfile(fai):
            entries = self._build_index()
            try:
                with open(fai, "w") as out:
                    for name, (length, offset, bases, width) in entries.items():
                        out.write(f"{name}\t{length}\t{offset}\t{bases}\t{width}\n")
            except OSError:
                pass
            return entries
        entries = {}
        with open(fai) as fh:
            for line in fh:
                f = line.rstrip("\n").split("\t")
                if len(f) >= 5:
                    entries[f[0]] = (int(f[1]), int(f[2]), int(f[3]), int(f[4]))
        return entries

    def _build_index(self):
        entries = {}
        name, length, offset, bases, width = None, 0, 0, 0, 0
        pos = 0
        with open(self._fasta_path, "rb") as fh:
            for raw in fh:
                if raw.startswith(b">"):
                    if name is not None:
                        entries[name] = (length, offset, bases, width)
                    name = raw[1:].split()[0].decode()
                    length, offset, bases, width = 0, pos + len(raw), 0, 0
                else:
                    stripped = raw.rstrip(b"\r\n")
                    if bases == 0 and stripped:
                        bases, width = len(stripped), len(raw)
                    length += len(stripped)
                pos += len(raw)
        if name is not None:
            entries[name] = (length, offset, bases, width)
        return entries

    # -- reference API -------------------------------------------------------------------------
    def close(self):
        if self._mm:
            self._mm.close()
        self._fh.close()

    @property
    def chroms(self):
        if self._chroms is None:
            self._chroms = sorted(self._index)  # genome/__init__.py:99
        return self._chroms

    @property
    def chrom_sizes(self):
        if self._chrom_sizes is None:
            self._chrom_sizes = {c: self._index[c][0] for c in self.chroms}
        return self._chrom_sizes

    def fetch_bytes(self, chrom, start, end):
        length, offset, bases, width = self._index[chrom]  # KeyError for unknown chromosomes
        start = max(int(start), 0)
        end = min(int(end), length)
        if end <= start:
            return b""
        if bases == 0:
            return b""
        a = offset + (start // bases) * width + start % bases
        b = offset + ((end - 1) // bases) * width + (end - 1) % bases + 1
        raw = self._mm[a:b]
        if width != bases:
            raw = raw.replace(b"\n", b"").replace(b"\r", b"")
        return raw

    def fetch_sequence(self, chrom, start, end):
        return self.fetch_bytes(chrom, start, end).decode("ascii")

    def random_sequences(self, n_times, length, max_n=0, random_seed=None):
        """Background sampling with the reference's exact RNG call sequence
        (genome/__init__.py:159-176): legacy `np.random.seed`, one `np.random.choice` of `n_times`
        chromosomes weighted by size, then per attempt `np.random.randint(size - length)`;
        a sample is accepted iff it holds at most `max_n` 'N'/'n'."""
        if random_seed is not None:
            np.random.seed(random_seed)
        sizes = self.chrom_sizes
        total = sum(sizes.values())
        weights = [sizes[c] / total for c in self.chroms]
        picks = np.random.choice(self.chroms, size=n_times, p=weights)
        got = 0
        attempt = 0
        while got < n_times:
            chrom = picks[attempt % n_times]
            start = np.random.randint(sizes[chrom] - length)
            seq = self.fetch_sequence(chrom, start, start + length)
            if seq.count("N") + seq.count("n") <= max_n:
                yield seq
                got += 1
            attempt += 1


def read_bg_freq(path):
    """`Base<TAB>Frequency` table for A, C, G, T (reference genome/__init__.py:223-260)."""
    freq = {}
    with open(path) as fh:
        for line in fh:
            f = line.split()
            if len(f) == 2 and f[0] in BASES:
                freq[f[0]] = float(f[1])
    if sorted(freq) != sorted(BASES):
        raise ValueError(f"invalid background frequency file: {path}")
    return freq
