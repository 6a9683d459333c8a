"""world_size-2 gloo test (CPU) of the multi-GPU host logic: shard by sequence, scan each shard
independently, gather on rank 0, restore the reference order -- must equal the unsharded scan.
The per-shard scanner here is the oracle (this test checks the plumbing, not the kernel)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import oracle
    from motifscan_b200 import shard
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)  # same data on every rank
        pwms = [np.around(rng.normal(0, 2, size=(4, int(rng.integers(4, 20)))), 5).tolist() for _ in range(7)]
        alphabet = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)
        seqs = [bytes(alphabet[rng.integers(0, 9, size=int(rng.integers(0, 400)))]).decode() for _ in range(41)]
        cutoffs = [0.35] * len(pwms)
        a, b = shard.region_block(len(seqs), world, rank)
        counts, seq_idx, start, score, strand = oracle.scan_arrays(pwms, cutoffs, seqs[a:b], 3)
        local = dict(counts=counts, motif=np.repeat(np.arange(len(pwms), dtype=np.int32), counts),
                     seq=seq_idx + a, start=start, score=score, strand=strand)
        merged = shard.gather_sites(local, len(pwms), dist=dist, dst=0)
        if rank == 0:
            full = oracle.scan_arrays(pwms, cutoffs, seqs, 3)
            ok = (np.array_equal(merged["counts"], full[0]) and np.array_equal(merged["seq"], full[1])
                  and np.array_equal(merged["start"], full[2])
                  and np.array_equal(merged["score"].view(np.uint64), full[3].view(np.uint64))
                  and np.array_equal(merged["strand"], full[4]) and int(full[0].sum()) > 50)
            q.put(bool(ok))
        else:
            assert merged is None
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_two_rank_gather_restores_reference_order():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_region_block_and_lpt():
    from motifscan_b200 import shard
    for n in (0, 1, 7, 50000):
        for world in (1, 2, 3, 8):
            blocks = [shard.region_block(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    from motifscan_b200.synth import HG19_NAMES, HG19_SIZES
    chunks = shard.genome_chunks(dict(zip(HG19_NAMES, HG19_SIZES)), 16 << 20, 29)
    assert sum(e - s for _, s, e, _ in chunks) == sum(HG19_SIZES)
    assert all(f - e <= 29 for _, _, e, f in chunks)
    owner, load = shard.assign_lpt([e - s for _, s, e, _ in chunks], 8)
    assert load.max() / load.mean() < 1.02   # chunked hg19 balances to within 2 % over 8 GPUs
    whole, wl = shard.assign_lpt(HG19_SIZES, 8)
    assert wl.max() / wl.mean() > load.max() / load.mean()
