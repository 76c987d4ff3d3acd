"""CPU tests of the multi-GPU steps 2-3 orchestration (ntjoin_b200.dist): the layout, the lock-step simulator and a
real world-size-2 gloo run, with numpy stand-ins for the device stages and the oracle as the single-process truth."""
import os
import sys

import numpy as np
import pytest
import torch

from ntjoin_b200 import synth
from ntjoin_b200.dist import Layout, merge_shards, run_lockstep, shard_ranges

from dist_ref_stages import NumpyA2AStages, NumpyDistStages

STAGES = {"allreduce": NumpyDistStages, "alltoall": NumpyA2AStages}

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WEIGHTS3 = [2.0, 2.0, 1.0]


def make_case(oracle, n_ref=2, bases=300_000, w=100, seed=7):
    """references + target (assembly order: references first, target last), sketched by the oracle"""
    rseq, roffs, _ = synth.make_reference(bases, n_chrom=5, dup_frac=0.05, seed=seed)
    asms = [(rseq, roffs)]
    for i in range(n_ref):
        t = synth.derive_target(rseq, roffs, min_len=3000, max_len=40000, seed=seed + 10 + i)
        asms.append((t[0], t[1]))
    asms = asms[1:] + asms[:1] if n_ref else asms            # keep the ancestor last (it plays the target)
    full = [oracle.sketch(s, o, 32, w) for s, o in asms]
    return asms, full


def shard_case(oracle, asms, world, w=100):
    """per rank: hashes / contigs tensors of its contiguous record ranges (oracle stands in for the GPU sketch)"""
    rr = shard_ranges([o for _, o in asms], world)
    hashes, contigs = [], []
    for r in range(world):
        hh, cc = [], []
        for (seq, offs), (c0, c1) in zip(asms, rr[r]):
            lo, hi = int(offs[c0]), int(offs[c1])
            m = oracle.sketch(seq[lo:hi], (offs[c0:c1 + 1] - offs[c0]).astype(np.uint64), 32, w)
            hh.append(torch.from_numpy(m["out_hash"].view(np.int64).copy()))
            cc.append(torch.from_numpy(m["contig"].astype(np.int32)))
        hashes.append(hh)
        contigs.append(cc)
    return hashes, contigs


def check_merged(merged, want):
    for a in range(len(want["uniq"])):
        assert np.array_equal(merged["uniq"][a].astype(bool), want["uniq"][a])
        assert np.array_equal(merged["keep"][a].astype(bool), want["keep"][a])
    assert np.array_equal(merged["vertices"], want["vertices"])
    assert np.array_equal(merged["edge_u"], want["edges"]["u"]) and np.array_equal(merged["edge_v"], want["edges"]["v"])
    assert np.array_equal(merged["support"], want["edges"]["support_mask"])
    assert np.array_equal(merged["weight"], want["edges"]["weight"])


def test_layout_offsets():
    lay = Layout([[3, 0], [2, 5], [0, 1]])
    assert lay.N == 11 and list(lay.asm_off) == [0, 5, 11]
    assert lay.goff.tolist() == [[0, 5], [3, 5], [5, 10]]


@pytest.mark.parametrize("mode", ["alltoall", "allreduce"])
@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_lockstep_matches_oracle(oracle, world, mode):
    asms, full = make_case(oracle)
    want = oracle.filter_and_edges([f["out_hash"] for f in full], [f["contig"] for f in full], WEIGHTS3)
    assert len(want["edges"]) > 100
    hashes, contigs = shard_case(oracle, asms, world)
    shards = run_lockstep([STAGES[mode]() for _ in range(world)], hashes, contigs, WEIGHTS3, torch.device("cpu"))
    check_merged(merge_shards([s.fetch() for s in shards]), want)


@pytest.mark.parametrize("mode", ["alltoall", "allreduce"])
def test_lockstep_empty_rank_and_no_survivors(oracle, mode):
    """more ranks than records on one side, and assemblies with nothing in common"""
    a = synth.make_reference(60_000, n_chrom=2, seed=3)
    b = synth.make_reference(60_000, n_chrom=2, seed=4)
    asms = [(a[0], a[1]), (b[0], b[1])]
    full = [oracle.sketch(s, o, 32, 50) for s, o in asms]
    want = oracle.filter_and_edges([f["out_hash"] for f in full], [f["contig"] for f in full], [1.0, 1.0])
    assert len(want["vertices"]) == 0
    hashes, contigs = shard_case(oracle, asms, 6, w=50)
    shards = run_lockstep([STAGES[mode]() for _ in range(6)], hashes, contigs, [1.0, 1.0], torch.device("cpu"))
    check_merged(merge_shards([s.fetch() for s in shards]), want)


@pytest.mark.parametrize("mode", ["alltoall", "allreduce"])
def test_lockstep_single_assembly_and_heavy_duplicates(oracle, mode):
    """n_asm = 1 (every unique hash survives; edges come from one assembly) and an assembly full of repeated segments"""
    rseq, roffs, _ = synth.make_reference(250_000, n_chrom=7, dup_frac=0.4, seed=21)
    one = [(rseq, roffs)]
    full = [oracle.sketch(rseq, roffs, 32, 40)]
    want = oracle.filter_and_edges([full[0]["out_hash"]], [full[0]["contig"]], [1.5])
    assert 0 < want["uniq"][0].sum() < len(full[0])          # duplicates were dropped
    for world in (2, 3):
        hashes, contigs = shard_case(oracle, one, world, w=40)
        shards = run_lockstep([STAGES[mode]() for _ in range(world)], hashes, contigs, [1.5], torch.device("cpu"))
        check_merged(merge_shards([s.fetch() for s in shards]), want)
    tseq, toffs, _ = synth.derive_target(rseq, roffs, min_len=2000, max_len=30000, seed=22)
    two = [(rseq, roffs), (tseq, toffs)]
    full = [oracle.sketch(s, o, 32, 40) for s, o in two]
    want = oracle.filter_and_edges([f["out_hash"] for f in full], [f["contig"] for f in full], [2.0, 1.0])
    hashes, contigs = shard_case(oracle, two, 4, w=40)
    shards = run_lockstep([STAGES[mode]() for _ in range(4)], hashes, contigs, [2.0, 1.0], torch.device("cpu"))
    check_merged(merge_shards([s.fetch() for s in shards]), want)


def _worker(rank, world, port, q, mode):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle_lib
    from ntjoin_b200.dist import TorchComm, distributed_filter_and_edges
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle_lib.Oracle()
    asms, full = make_case(orc)
    hashes, contigs = shard_case(orc, asms, world)
    shard = distributed_filter_and_edges(STAGES[mode](), hashes[rank], contigs[rank], WEIGHTS3, TorchComm(torch.device("cpu")))
    gathered = [None] * world
    dist.all_gather_object(gathered, shard.fetch())
    ok = True
    if rank == 0:
        want = orc.filter_and_edges([f["out_hash"] for f in full], [f["contig"] for f in full], WEIGHTS3)
        try:
            check_merged(merge_shards(gathered), want)
        except AssertionError:
            ok = False
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["alltoall", "allreduce"])
def test_two_rank_gloo_distributed_filter(mode):
    """world_size 2 over gloo: the exchanges of either formulation reproduce the single-process steps 2-3"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port + (7 if mode == "alltoall" else 0), q, mode)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in out) == [0, 1] and all(ok for _, ok in out)


def _tsv_worker(rank, world, port, q, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import subprocess
    import torch.distributed as dist
    import oracle_lib
    from ntjoin_b200.dist import gather_and_write_tsv
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle_lib.Oracle()
    fa = os.path.join(tmp, "asm.fa")
    names, seq, offs = oracle_lib.read_fasta(fa)
    c0, c1 = shard_ranges([offs], world)[rank][0]
    lo, hi = int(offs[c0]), int(offs[c1])
    m = orc.sketch(seq[lo:hi], (offs[c0:c1 + 1] - offs[c0]).astype(np.uint64), 32, 50)     # this rank's record range
    out = os.path.join(tmp, "asm.fa.k32.w50.tsv")
    gather_and_write_tsv(out, m["out_hash"], m["pos"], m["contig"], m["forward"], c0, names, 32, fasta_path=fa)
    ok = True
    if rank == 0:
        want = subprocess.check_output([oracle_lib.CLI, "--seq", "--long", "--pos", "-k", "32", "-w", "50", fa])
        ok = open(out, "rb").read() == want
    q.put((rank, ok))
    dist.destroy_process_group()


def test_two_rank_gloo_tsv_of_sharded_sketch(tmp_path):
    """seam S2 with two ranks: each sketches its record range, rank 0 writes the assembly's one TSV, byte-equal to a
    single-process `indexlr --seq --long --pos`"""
    import torch.multiprocessing as mp
    rseq, roffs, rnames = synth.make_reference(400_000, n_chrom=3, n_frac=0.01, seed=31)
    seq, offs, names = synth.derive_target(rseq, roffs, min_len=500, max_len=30_000, seed=32)
    synth.write_fasta(str(tmp_path / "asm.fa"), seq, offs, names, width=70)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_tsv_worker, args=(r, 2, port, q, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in out) == [0, 1] and all(ok for _, ok in out)
