"""Deterministic synthetic workloads of the shapes BASELINE.json names.
Everything is generated on the box that runs the benchmark; nothing large lives in the repo."""
import numpy as np

from .motif.matrix import pfm_to_pwm

BG = dict(A=0.295, C=0.205, G=0.205, T=0.295)
BG_P = np.array([BG[b] for b in "ACGT"])

# hg19 chromosome lengths chr1..22, X, Y, M
HG19_SIZES = [249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663,
              146364022, 141213431, 135534747, 135006516, 133851895, 115169878, 107349540,
              102531392, 90354753, 81195210, 78077248, 59128983, 63025520, 48129895, 51304566,
              155270560, 59373566, 16571]
HG19_NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY", "chrM"]


def motif_set(n=750, seed=2020, lmin=6, lmax=30, n_long=0):
    """JASPAR-shaped PFMs pushed through the reference's PFM -> PPM(pseudo 0.001) -> PWM(bg,
    round 5) rules.  Returns (pfms, pwms, ids).  The last `n_long` motifs are 33-40 columns long
    (composite / dimer matrices): longer than the 32 columns the prefilters hold."""
    rng = np.random.default_rng(seed)
    pfms, pwms, ids = [], [], []
    for k in range(n):
        L = int(rng.integers(lmin, lmax + 1))
        if k >= n - n_long:
            L = 33 + (k - (n - n_long)) % 8
        depth = float(np.exp(rng.uniform(np.log(20), np.log(5000))))
        pfm = np.round(depth * rng.dirichlet([0.3] * 4, size=L).T).astype(np.int64)
        empty = pfm.sum(axis=0) == 0
        pfm[0, empty] = 1
        pfms.append(pfm)
        pwms.append(pfm_to_pwm(pfm, BG))
        ids.append(f"MA{k:04d}.1")
    return pfms, pwms, ids


def random_bases(rng, n, lower_frac=0.5, mean_run=300):
    """iid A/C/G/T bytes with the synthetic-genome composition, soft-masked in runs."""
    s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.choice(4, size=n, p=BG_P)]
    if lower_frac > 0 and n > 0:
        # alternate upper / lower runs with geometric lengths
        n_runs = max(int(2 * n / mean_run) + 2, 2)
        lens = rng.geometric(1.0 / mean_run, size=n_runs)
        edges = np.minimum(np.cumsum(lens), n)
        state = np.zeros(n, dtype=bool)
        lower = rng.random() < lower_frac
        a = 0
        for e in edges:
            if lower:
                state[a:e] = True
            lower = not lower
            a = e
            if a >= n:
                break
        s = s.copy()
        s[state] |= 0x20
    return s


def peak_set(n_regions=50000, width=1000, seed=50, n_frac=0.01):
    """`n_regions` fixed-width peak sequences as one uint8 blob + offsets.  A fraction `n_frac`
    of the peaks carries a short run of N (peaks are placed on non-gap sequence, but assembly
    gaps inside peaks do occur)."""
    rng = np.random.default_rng(seed)
    blob = random_bases(rng, n_regions * width)
    k = int(n_regions * n_frac)
    if k:
        which = rng.choice(n_regions, size=k, replace=False)
        for r in which:
            a = int(rng.integers(0, width - 1))
            b = min(width, a + int(rng.integers(1, 50)))
            blob[r * width + a:r * width + b] = ord("N")
    seq_off = np.arange(n_regions + 1, dtype=np.int64) * width
    return blob, seq_off


def background_samples(n, length, seed=1):
    """`n` background sequences of `length` bases without N (the `--max-n 0` sampling of
    motif --build on an iid genome), as blob + offsets."""
    rng = np.random.default_rng(seed)
    blob = random_bases(rng, n * length, lower_frac=0.0)
    return blob, np.arange(n + 1, dtype=np.int64) * length


def genome_chunk(n_bases, seed, n_block=None):
    """A chunk of hg19-shaped synthetic chromosome: iid bases, soft-masked runs, optional N
    block (start, length)."""
    rng = np.random.default_rng(seed)
    s = random_bases(rng, n_bases)
    if n_block is not None:
        a, ln = n_block
        s[a:a + ln] = ord("N")
    return s


def pack_ascii(blob):
    """numpy restatement of the library's packed layout for ONE sequence (include/msb200.h,
    msb_seqs_from_packed): (codes uint32[2 B], nmask uint32[B]), B = ceil(len / 32).  Test / bench
    infrastructure: the product encodes on the device (encode_pack_kernel)."""
    blob = np.asarray(blob, dtype=np.uint8)
    n = blob.size
    n_blocks = (n + 31) // 32
    lut = np.full(256, 4, dtype=np.uint8)
    for k, ch in enumerate(b"ACGT"):
        lut[ch] = k
        lut[ch | 0x20] = k
    c = np.zeros(n_blocks * 32, dtype=np.uint8)
    c[:n] = lut[blob]
    isn = c == 4
    isn[n:] = False
    c[c == 4] = 0
    c[n:] = 0
    sh2 = (2 * np.arange(16, dtype=np.uint64))[None, :]
    codes = (c.reshape(-1, 16).astype(np.uint64) << sh2).sum(axis=1).astype(np.uint32)
    sh1 = np.arange(32, dtype=np.uint64)[None, :]
    nmask = (isn.reshape(-1, 32).astype(np.uint64) << sh1).sum(axis=1).astype(np.uint32)
    return codes, nmask


def packed_genome(sizes, names=None, seed=19, n_frac=0.07, pinned=False):
    """A small synthetic `genome.PackedGenome` for tests: iid bases of the synthetic composition, 10 kb-scaled
    N runs at both chromosome ends and one interior N block per chromosome (about `n_frac` of it)."""
    from .genome import PackedGenome
    rng = np.random.default_rng(seed)
    names = names or [f"chr{i + 1:02d}" for i in range(len(sizes))]
    order = sorted(range(len(sizes)), key=lambda i: names[i])
    chroms = [names[i] for i in order]
    cs, ms = [], []
    for i in order:
        n = int(sizes[i])
        s = random_bases(rng, n)
        if n_frac > 0 and n >= 200:
            edge = max(1, min(10000, n // 50))
            s[:edge] = ord("N")
            s[n - edge:] = ord("N")
            ln = int(n * n_frac)
            a = int(rng.integers(edge, max(n - edge - ln, edge + 1)))
            s[a:a + ln] = ord("N")
        c, m = pack_ascii(s)
        cs.append(c)
        ms.append(m)
    codes = np.concatenate(cs) if cs else np.zeros(0, np.uint32)
    nmask = np.concatenate(ms) if ms else np.zeros(0, np.uint32)
    if pinned:
        pc, pm, pin = PackedGenome._alloc(nmask.size, True)
        pc[:], pm[:] = codes, nmask
        codes, nmask = pc, pm
    else:
        pin = None
    return PackedGenome(chroms, {names[i]: int(sizes[i]) for i in order}, codes, nmask, name="synthetic", _pin=pin)
