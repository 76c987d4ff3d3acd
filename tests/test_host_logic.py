"""CPU tests of host-side logic: synthetic generators, record sharding, the N>1 exchange (gloo, world 2)."""
import os
import sys

import numpy as np
import pytest

from ntjoin_b200 import synth
from ntjoin_b200.dist import shard_ranges

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synth_deterministic_and_shaped():
    a = synth.make_reference(200_000, n_chrom=4, dup_frac=0.05, n_frac=0.01)
    b = synth.make_reference(200_000, n_chrom=4, dup_frac=0.05, n_frac=0.01)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert set(np.unique(a[0])) <= set(b"ACGTN") and (a[0] == ord("N")).sum() > 0
    t = synth.derive_target(a[0], a[1], min_len=2000, max_len=30000)
    assert int(t[1][-1]) == len(t[0]) == len(a[0]) and len(t[2]) == len(t[1]) - 1
    assert np.array_equal(synth.revcomp(synth.revcomp(a[0][:1000])), a[0][:1000])


def test_shard_ranges_partition():
    rng = np.random.default_rng(1)
    offs = [np.concatenate([[0], np.cumsum(rng.integers(1, 1000, n))]).astype(np.uint64) for n in (24, 700)]
    for world in (1, 2, 3, 8):
        rr = shard_ranges(offs, world)
        assert len(rr) == world
        for a in range(2):
            assert rr[0][a][0] == 0 and rr[-1][a][1] == len(offs[a]) - 1
            for r in range(world - 1):
                assert rr[r][a][1] == rr[r + 1][a][0]          # contiguous, no gap, no overlap
        loads = [sum(int(offs[a][c1] - offs[a][c0]) for a, (c0, c1) in enumerate(rr[r])) for r in range(world)]
        assert sum(loads) == sum(int(o[-1]) for o in offs)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle_lib
    from ntjoin_b200.dist import all_gather_minimizers
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle_lib.Oracle()
    rseq, roffs, _ = synth.make_reference(400_000, n_chrom=5, dup_frac=0.05)
    tseq, toffs, _ = synth.derive_target(rseq, roffs, min_len=3000, max_len=40000)
    asms = [(rseq, roffs), (tseq, toffs)]
    mine = shard_ranges([o for _, o in asms], world)[rank]
    hashes, contigs = [], []
    for (seq, offs), (c0, c1) in zip(asms, mine):
        lo, hi = int(offs[c0]), int(offs[c1])
        m = orc.sketch(seq[lo:hi], (offs[c0:c1 + 1] - offs[c0]).astype(np.uint64), 32, 100)   # oracle stands in for the GPU sketch
        hh, cc = all_gather_minimizers(torch.from_numpy(m["out_hash"].view(np.int64).copy()),
                                       torch.from_numpy(m["contig"].astype(np.int32)), c0)
        hashes.append(hh.numpy().view(np.uint64))
        contigs.append(cc.numpy().astype(np.uint32))
    res = orc.filter_and_edges(hashes, contigs, [2.0, 1.0])
    full = [orc.sketch(s, o, 32, 100) for s, o in asms]
    ok = all(np.array_equal(h, f["out_hash"]) and np.array_equal(c, f["contig"]) for h, c, f in zip(hashes, contigs, full))
    want = orc.filter_and_edges([f["out_hash"] for f in full], [f["contig"] for f in full], [2.0, 1.0])
    ok = ok and np.array_equal(res["edges"], want["edges"]) and np.array_equal(res["vertices"], want["vertices"])
    q.put((rank, bool(ok), len(want["edges"])))
    dist.destroy_process_group()


def test_two_rank_exchange_gloo():
    """world_size 2 over gloo: sharded sketches + one all-gather reproduce the single-process result"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _, _ in out) == [0, 1]
    assert all(ok for _, ok, _ in out) and out[0][2] > 100
