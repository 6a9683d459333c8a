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


class DeviceGenome:
    """The chromosomes of a genome encoded once (2-bit codes + N mask, 0.375 B/bp) and kept resident
    in one GPU's HBM.

    The reference fetches every region's sequence from the FASTA through pysam
    (scanner.py:76-87 -> genome/__init__.py:135) and hands strings to the extension; here a
    `Scanner` built on a DeviceGenome ships (chromosome, start, end) descriptors and the device
    cuts the packed windows out of the resident copy (`msb_seqs_extract`), and a genome-wide scan
    walks position ranges of the resident chromosomes (`msb_scan_ranges`).  Everything else
    (`chroms`, `chrom_sizes`, `fetch_sequence`, `random_sequences`, `bg_freq`, ...) is the
    wrapped host genome's."""

    def __init__(self, genome, ctx=None, device=0):
        from . import engine
        self.genome = genome
        self.ctx = ctx or engine.default_context(device)
        self.chroms = list(genome.chroms)
        self.chrom_sizes = {c: int(genome.chrom_sizes[c]) for c in self.chroms}
        self.chrom_index = {c: i for i, c in enumerate(self.chroms)}
        off = np.zeros(len(self.chroms) + 1, dtype=np.int64)
        np.cumsum([self.chrom_sizes[c] for c in self.chroms], out=off[1:])
        blob = np.empty(int(off[-1]), dtype=np.uint8)
        fetch = getattr(genome, "fetch_bytes", None)
        for i, c in enumerate(self.chroms):
            raw = fetch(c, 0, self.chrom_sizes[c]) if fetch else genome.fetch_sequence(c, 0, self.chrom_sizes[c]).encode()
            if len(raw) != off[i + 1] - off[i]:
                raise ValueError(f"chromosome {c}: fetched {len(raw)} bases, index says {off[i + 1] - off[i]}")
            blob[off[i]:off[i + 1]] = np.frombuffer(raw, dtype=np.uint8)
        self.seqs = engine.SequenceSet(self.ctx, blob=blob, seq_off=off)

    def __getattr__(self, name):
        # only reached for attributes this class does not define: delegate to the host genome
        return getattr(self.__dict__["genome"], name)

    def extract(self, chroms, starts, ends):
        """Device sequence set of the intervals [start, end) (end clipped at the chromosome size);
        an unknown chromosome raises KeyError like `Genome.fetch_sequence`."""
        idx = np.fromiter((self.chrom_index[c] for c in chroms), dtype=np.int32, count=len(chroms))
        return self.seqs.extract(idx, starts, ends)

    def random_windows(self, n_times, length, max_n=0, random_seed=None):
        """The background sampler (reference genome/__init__.py:159-176) without fetching the samples:
        the same legacy-numpy RNG calls in the same order -- `np.random.seed`, one `np.random.choice`
        of `n_times` chromosomes weighted by size, then one `np.random.randint(size - length)` per
        attempt (drawn as arrays: the legacy generator gives the same stream for an array of bounds
        as for the scalar calls, checked in tests) -- and the same acceptance rule, with the N count of
        an attempt taken from the resident N mask.  Only attempts that hold SOME non-ACGT base and
        could still pass (`max_n` > 0, or IUPAC codes, which the reference does not count) are fetched
        and counted on the host.  Returns (chromosome indices, starts) of the accepted samples in
        order; the global RNG is left exactly where the reference leaves it."""
        if random_seed is not None:
            np.random.seed(random_seed)
        sizes = np.array([self.chrom_sizes[c] for c in self.chroms], dtype=np.int64)
        weights = [s / sizes.sum() for s in sizes.tolist()]          # genome/__init__.py:163-165
        picks = np.random.choice(self.chroms, size=n_times, p=weights)
        pick_idx = np.fromiter((self.chrom_index[c] for c in picks), dtype=np.int32, count=n_times)
        acc_idx, acc_start = [], []
        got, attempt = 0, 0
        while got < n_times:
            m = max(1024, min(2 * (n_times - got), 1 << 22))
            state = np.random.get_state()
            idx = pick_idx[(attempt + np.arange(m)) % n_times]
            starts = np.random.randint(sizes[idx] - length)
            ncount = self.seqs.window_ncount(idx, starts, length)
            ok = ncount <= max_n
            for k in np.flatnonzero(~ok):                            # may still pass by the reference's own count
                seq = self.genome.fetch_sequence(self.chroms[idx[k]], int(starts[k]), int(starts[k]) + length)
                ok[k] = seq.count("N") + seq.count("n") <= max_n
            cum = np.cumsum(ok)
            need = n_times - got
            if cum[-1] >= need:
                used = int(np.searchsorted(cum, need)) + 1            # attempts of this batch the reference makes
                np.random.set_state(state)
                np.random.randint(sizes[idx[:used]] - length)         # consume exactly those draws
                ok[used:] = False
                m = used
            acc_idx.append(idx[:m][ok[:m]])
            acc_start.append(starts[:m][ok[:m]])
            got += int(ok[:m].sum())
            attempt += m
        return np.concatenate(acc_idx), np.concatenate(acc_start).astype(np.int64)

    def random_sequence_set(self, n_times, length, max_n=0, random_seed=None):
        """`random_sequences` as a device-resident sequence set: sampled with the reference's RNG and cut
        out of the resident genome, no strings."""
        idx, starts = self.random_windows(n_times, length, max_n, random_seed)
        return self.seqs.extract(idx, starts, starts + length)

    def close(self):
        self.seqs.close()


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
