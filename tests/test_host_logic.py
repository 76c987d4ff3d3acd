"""CPU tests of host-side logic: synthetic generators and record sharding (the N>1 exchanges: test_dist_logic.py)."""
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
        total = sum(int(o[-1]) for o in offs)
        assert sum(loads) == total
        assert max(loads) <= total / world + 1000               # within one (finest) record of perfect balance


def test_shard_ranges_compensates_coarse_assembly():
    """chromosome-scale records in one assembly are evened out by the contigs of the other"""
    big = np.concatenate([[0], np.cumsum([248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133])]).astype(np.uint64) * 1_000_000
    rng = np.random.default_rng(2)
    small = np.concatenate([[0], np.cumsum(rng.integers(50_000, 5_000_000, 900))]).astype(np.uint64)
    total = int(big[-1] + small[-1])
    for world in (2, 4, 8):
        rr = shard_ranges([big, small], world)
        loads = [int(big[rr[r][0][1]] - big[rr[r][0][0]]) + int(small[rr[r][1][1]] - small[rr[r][1][0]]) for r in range(world)]
        assert sum(loads) == total and max(loads) <= total / world * 1.02


def _indexlr_module():
    import importlib.machinery
    import importlib.util
    path = os.path.join(ROOT, "bin", "indexlr")
    loader = importlib.machinery.SourceFileLoader("indexlr_cli", path)
    spec = importlib.util.spec_from_loader("indexlr_cli", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def test_indexlr_cli_argv_forms():
    """bin/indexlr parses both argv forms the reference uses (ntJoin:205, bin/ntjoin_utils.py:197-198) and rejects the rest"""
    cli = _indexlr_module()
    a = cli.parse("--seq --long --pos -k 32 -w 1000 -t 4 asm.fa".split())
    b = cli.parse("asm.fa --seq --long --pos -k32 -w1000 -t4 -o asm.fa.k32.w1000.tsv".split())
    for o in (a, b):
        assert (o["k"], o["w"], o["t"], o["pos"], o["seq"], o["strand"], o["files"]) == (32, 1000, 4, True, True, False, ["asm.fa"])
    assert a["out"] == "-" and b["out"] == "asm.fa.k32.w1000.tsv" and a["canonical"] == "sum"
    assert cli.parse("-k 15 -w 10 --canonical min x.fa y.fa".split())["files"] == ["x.fa", "y.fa"]
    for bad in (["-k", "32", "x.fa"], ["-k", "32", "-w", "10"], ["-k", "0", "-w", "5", "x.fa"], ["--frobnicate", "-k", "3", "-w", "3", "x.fa"],
                ["-k"]):
        with pytest.raises(SystemExit) as ei:
            cli.parse(bad)
        assert ei.value.code not in (0, None)


def test_mxlists_pickle_as_plain_lists():
    """bin/ntjoin.py:173-174 pickles the Ntjoin object (list_mxs included) into multiprocessing.Pool workers when
    assemble_t > 1; the engine handles hidden in the drop-in's lists (ctypes pointers) must not break that"""
    import ctypes
    import pickle
    from ntjoin_b200.dropin import MxLists

    class FakeSketch:
        def __init__(self):
            self.handle = ctypes.c_void_p(1234)      # what makes a plain pickle fail

    lists = MxLists([["1", "2"], [], ["3"]])
    lists._sketch, lists._mask, lists._asm_index = FakeSketch(), np.array([True, False]), 0
    with pytest.raises((ValueError, TypeError)):
        pickle.dumps(lists._sketch.handle)
    back = pickle.loads(pickle.dumps({"asm.tsv": lists}))
    assert type(back["asm.tsv"]) is list and back["asm.tsv"] == [["1", "2"], [], ["3"]]


def test_bucket_ownership_formulas_agree_for_every_world_size():
    """csrc/p2p.cu: the kernels route a record to owner(b) = (b * world) >> B, the host gives rank r the buckets
    [fb[r], fb[r + 1]) with fb[r] = ceil(r * 2^B / world).  The two must describe the same partition -- monotone in the
    hash, every bucket owned once, no rank empty -- for every world size up to 16 and every bucket count (the GPU tests
    run worlds 1, 2, 3, 4, 5 and 8)."""
    for world in range(1, 17):
        for B in range(6, 21):
            n_buckets = 1 << B
            fb = [((r << B) + world - 1) // world for r in range(world + 1)]
            assert fb[0] == 0 and fb[world] == n_buckets
            owner = (np.arange(n_buckets, dtype=np.int64) * world) >> B
            assert (np.diff(owner) >= 0).all()
            for r in range(world):
                seg = owner[fb[r]:fb[r + 1]]
                assert len(seg) > 0 and (seg == r).all(), (world, B, r)
