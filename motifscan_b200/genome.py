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

    def count_n(self, chroms, starts, length):
        """`seq.count('N') + seq.count('n')` (genome/__init__.py:172-175) of the windows [start, start + length)
        of the named chromosomes, for many windows at once: one gather from the memory-mapped FASTA instead of
        one fetch per window (windows are clipped at the chromosome end like fetch)."""
        starts = np.asarray(starts, dtype=np.int64)
        out = np.zeros(len(starts), dtype=np.int64)
        if len(starts) == 0 or length <= 0:
            return out
        data = np.frombuffer(self._mm, dtype=np.uint8)
        meta = np.array([self._index[c] for c in chroms], dtype=np.int64).reshape(-1, 4)   # length, offset, bases, width
        size, offset, bases, width = meta[:, 0], meta[:, 1], np.maximum(meta[:, 2], 1), meta[:, 3]
        pos = starts[:, None] + np.arange(length, dtype=np.int64)[None, :]
        inside = (pos >= 0) & (pos < size[:, None])
        pos = np.where(inside, pos, 0)
        at = offset[:, None] + (pos // bases[:, None]) * width[:, None] + pos % bases[:, None]
        byte = data[np.minimum(at, data.size - 1)]
        return (((byte == ord("N")) | (byte == ord("n"))) & inside).sum(axis=1)

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


class PackedGenome:
    """A genome in the library's packed representation on the HOST: 2-bit codes + N mask of every
    chromosome (0.375 B/bp; hg19 = 1.16 GB), each chromosome padded to 32 bases, planes concatenated in
    `chroms` order, in page-locked memory when a CUDA device is there.

    This is the host-side source of every device copy: `DeviceGenome(PackedGenome)` uploads the planes
    as they are (`msb_seqs_from_packed`), and a multi-GPU genome scan gives each GPU one contiguous
    block range of them (`genome_scan.plan_shares`).  Built once per genome from the FASTA (encoded on a
    GPU, chromosome by chromosome) and cached next to it with `save` / `load`, the way the reference
    keeps a `.fai` next to its FASTA (genome/__init__.py:61-95).  Soft-masking (lower case) is not kept:
    scoring never sees it (cscore.c:91-108 folds case)."""

    def __init__(self, chroms, chrom_sizes, codes, nmask, name="packed", genome=None, _pin=None):
        self.name = name
        self.chroms = list(chroms)
        self.chrom_sizes = {c: int(chrom_sizes[c]) for c in self.chroms}
        self.chrom_index = {c: i for i, c in enumerate(self.chroms)}
        blocks = np.array([(self.chrom_sizes[c] + 31) // 32 for c in self.chroms], dtype=np.int64)
        self.block_off = np.zeros(len(self.chroms) + 1, dtype=np.int64)
        np.cumsum(blocks, out=self.block_off[1:])
        self.n_blocks = int(self.block_off[-1])
        if codes.size != 2 * self.n_blocks or nmask.size != self.n_blocks:
            raise ValueError("packed planes do not match the chromosome sizes")
        self.codes, self.nmask = codes, nmask
        self.genome = genome          # optional host genome the planes came from (bg_freq, strings with case)
        self._pin = _pin              # keeps the pinned allocations alive

    @staticmethod
    def _alloc(n_blocks, pinned=True):
        """(codes, nmask, keepalive): page-locked when the CUDA library can allocate it."""
        if pinned:
            try:
                from . import engine
                a, b = engine.PinnedArray(max(8 * n_blocks, 8)), engine.PinnedArray(max(4 * n_blocks, 4))
                return a.array[:8 * n_blocks].view(np.uint32), b.array[:4 * n_blocks].view(np.uint32), (a, b)
            except Exception:
                pass
        return np.zeros(2 * n_blocks, dtype=np.uint32), np.zeros(n_blocks, dtype=np.uint32), None

    @classmethod
    def from_genome(cls, genome, ctx=None, pinned=True):
        """Encode a host genome (anything with chroms / chrom_sizes / fetch_bytes or fetch_sequence) on the
        device, one chromosome at a time, and keep the packed planes on the host."""
        from . import engine
        ctx = ctx or engine.default_context(0)
        chroms = list(genome.chroms)
        sizes = {c: int(genome.chrom_sizes[c]) for c in chroms}
        n_blocks = sum((sizes[c] + 31) // 32 for c in chroms)
        codes, nmask, pin = cls._alloc(n_blocks, pinned)
        fetch = getattr(genome, "fetch_bytes", None)
        at = 0
        for c in chroms:
            raw = fetch(c, 0, sizes[c]) if fetch else genome.fetch_sequence(c, 0, sizes[c]).encode()
            if len(raw) != sizes[c]:
                raise ValueError(f"chromosome {c}: fetched {len(raw)} bases, index says {sizes[c]}")
            sset = engine.SequenceSet(ctx, blob=np.frombuffer(raw, dtype=np.uint8), seq_off=np.array([0, sizes[c]], dtype=np.int64))
            try:
                cw, mw = sset.to_packed()
            finally:
                sset.close()
            nb = (sizes[c] + 31) // 32
            codes[2 * at:2 * (at + nb)] = cw
            nmask[at:at + nb] = mw
            at += nb
        return cls(chroms, sizes, codes, nmask, name=getattr(genome, "name", "packed"), genome=genome, _pin=pin)

    def save(self, prefix):
        """`<prefix>.codes.npy`, `<prefix>.nmask.npy`, `<prefix>.chroms.tsv`: the packed-genome cache."""
        np.save(prefix + ".codes.npy", np.asarray(self.codes))
        np.save(prefix + ".nmask.npy", np.asarray(self.nmask))
        with open(prefix + ".chroms.tsv", "w") as out:
            for c in self.chroms:
                out.write(f"{c}\t{self.chrom_sizes[c]}\n")

    @classmethod
    def load(cls, prefix, genome=None, pinned=True):
        chroms, sizes = [], {}
        with open(prefix + ".chroms.tsv") as fh:
            for line in fh:
                c, n = line.rstrip("\n").split("\t")
                chroms.append(c)
                sizes[c] = int(n)
        n_blocks = sum((sizes[c] + 31) // 32 for c in chroms)
        codes, nmask, pin = cls._alloc(n_blocks, pinned)
        codes[:] = np.load(prefix + ".codes.npy", mmap_mode="r")
        nmask[:] = np.load(prefix + ".nmask.npy", mmap_mode="r")
        return cls(chroms, sizes, codes, nmask, name=os.path.basename(prefix), genome=genome, _pin=pin)

    def __getattr__(self, name):
        g = self.__dict__.get("genome")
        if g is None:
            raise AttributeError(name)
        return getattr(g, name)

    def work_prefix(self):
        """Cumulative per-block scan cost (B + 1 entries): the number of A/C/G/T bases of a block plus a small
        constant -- windows inside N runs are dropped by whole position tiles, so the non-N bases are what a
        share costs.  `genome_scan.plan_units` cuts shares at equal increments of it."""
        w = self.__dict__.get("_work_prefix")
        if w is None:
            acgt = 32 - np.bitwise_count(np.asarray(self.nmask)).astype(np.int64)
            # the last block of a chromosome holds fewer than 32 bases; the constant stands for the per-tile overhead
            w = np.zeros(self.n_blocks + 1, dtype=np.int64)
            np.cumsum(acgt + 1, out=w[1:])
            self.__dict__["_work_prefix"] = w
        return w

    def planes(self, block0, block1):
        """Views of the planes of blocks [block0, block1) (no copy: they go to the device as they are)."""
        return self.codes[2 * block0:2 * block1], self.nmask[block0:block1]

    def fetch_bytes(self, chrom, start, end):
        """Bases [start, end): the source genome's own bytes when there is one (case and IUPAC codes as in
        the FASTA, which the background sampler's N count looks at, genome/__init__.py:172-175), else
        decoded from the planes."""
        g = self.__dict__.get("genome")
        if g is not None:
            fetch = getattr(g, "fetch_bytes", None)
            return fetch(chrom, start, end) if fetch else g.fetch_sequence(chrom, start, end).encode()
        return self.decode_bytes(chrom, start, end)

    def decode_bytes(self, chrom, start, end):
        """Bases [start, end) of a chromosome decoded from the planes to upper-case ASCII, non-ACGT as `N`
        (`end` clipped at the chromosome size, like pysam's fetch)."""
        size = self.chrom_sizes[chrom]      # KeyError for unknown chromosomes
        start, end = max(int(start), 0), min(int(end), size)
        if end <= start:
            return b""
        b0 = int(self.block_off[self.chrom_index[chrom]])
        w0, w1 = start // 16, (end + 15) // 16
        cw = np.asarray(self.codes[2 * b0 + w0:2 * b0 + w1])
        two = ((cw[:, None] >> (2 * np.arange(16, dtype=np.uint32))[None, :]) & 3).astype(np.uint8).ravel()
        m0, m1 = start // 32, (end + 31) // 32
        mw = np.asarray(self.nmask[b0 + m0:b0 + m1])
        isn = ((mw[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(bool).ravel()
        out = np.frombuffer(b"ACGT", dtype=np.uint8)[two[start - 16 * w0:end - 16 * w0]]
        out[isn[start - 32 * m0:end - 32 * m0]] = ord("N")
        return out.tobytes()

    def fetch_sequence(self, chrom, start, end):
        return self.fetch_bytes(chrom, start, end).decode("ascii")

    def count_n(self, chroms, starts, length):
        """N / n count of many windows (the background sampler's acceptance test): the source genome's when there
        is one; the planes' own mask otherwise (every masked base decodes to `N`)."""
        g = self.__dict__.get("genome")
        if g is not None and hasattr(g, "count_n"):
            return g.count_n(chroms, starts, length)
        starts = np.asarray(starts, dtype=np.int64)
        cidx = np.array([self.chrom_index[c] for c in chroms], dtype=np.int64)
        size = np.array([self.chrom_sizes[c] for c in chroms], dtype=np.int64)
        pos = starts[:, None] + np.arange(length, dtype=np.int64)[None, :]
        inside = (pos >= 0) & (pos < size[:, None])
        pos = np.where(inside, pos, 0)
        word = np.asarray(self.nmask)[self.block_off[cidx][:, None] + pos // 32]
        return ((((word >> (pos % 32).astype(np.uint32)) & 1) == 1) & inside).sum(axis=1)

    def close(self):
        self._pin = None


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
        if isinstance(genome, PackedGenome):    # already packed on the host: the planes go up as they are
            self.seqs = engine.SequenceSet.from_packed(self.ctx, [self.chrom_sizes[c] for c in self.chroms],
                                                       genome.codes, genome.nmask)
            return
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
            again = np.flatnonzero(~ok)                              # may still pass by the reference's own count:
            if len(again):                                           # it counts 'N' and 'n' only, the mask every non-ACGT byte
                count_n = getattr(self.genome, "count_n", None)
                if count_n is not None:
                    ok[again] = count_n([self.chroms[i] for i in idx[again]], starts[again], length) <= max_n
                else:
                    for k in again:
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
