"""GPU parity of the multi-GPU steps 2-3 (mxe_dist_* stages through the C ABI).

* lock-step: `world` simulated ranks on ONE device run the real device stages with the collectives done in place --
  sharded CUDA sketches + distributed steps 2-3 must equal the oracle and the single-GPU engine result;
* nccl: two processes on two devices (skipped on a single-GPU box) run the same over torch.distributed / NCCL.
"""
import os
import sys

import numpy as np
import pytest
import torch

from ntjoin_b200 import synth
from ntjoin_b200.dist import DeviceArray, merge_shards, run_lockstep, shard_ranges

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(n_ref, bases, seed=11, n_frac=0.003):
    rseq, roffs, _ = synth.make_reference(bases, n_chrom=6, dup_frac=0.03, n_frac=n_frac, seed=seed)
    asms = [synth.derive_target(rseq, roffs, seed=seed + 1 + i, min_len=4000, max_len=200_000, sub_rate=0.002)[:2] for i in range(n_ref)]
    return asms + [(rseq, roffs)]


def _rank_inputs(eng, asms, world, k, w, dev):
    """CUDA sketches of every rank's contiguous record range -> per rank (hashes, contigs) device tensors"""
    rr = shard_ranges([o for _, o in asms], world)
    hashes, contigs, keep = [], [], []
    for r in range(world):
        hh, cc = [], []
        for (seq, offs), (c0, c1) in zip(asms, rr[r]):
            lo, hi = int(offs[c0]), int(offs[c1])
            sk = eng.sketch_buffers(seq[lo:hi], (offs[c0:c1 + 1] - offs[c0]).astype(np.uint64), k, w)
            n, ph, _pp, pc = sk.device_pointers()
            hh.append(torch.as_tensor(DeviceArray(ph, n, "<i8"), device=dev).clone() if n else torch.empty(0, dtype=torch.int64, device=dev))
            cc.append(torch.as_tensor(DeviceArray(pc, n, "<i4"), device=dev).clone() if n else torch.empty(0, dtype=torch.int32, device=dev))
            keep.append(sk)
        hashes.append(hh)
        contigs.append(cc)
    return hashes, contigs, keep


def _check(merged, want):
    for a in range(len(want["uniq"])):
        np.testing.assert_array_equal(merged["uniq"][a].astype(bool), want["uniq"][a])
        np.testing.assert_array_equal(merged["keep"][a].astype(bool), want["keep"][a])
    np.testing.assert_array_equal(merged["vertices"], want["vertices"])
    np.testing.assert_array_equal(merged["edge_u"], want["edges"]["u"])
    np.testing.assert_array_equal(merged["edge_v"], want["edges"]["v"])
    np.testing.assert_array_equal(merged["support"], want["edges"]["support_mask"])
    np.testing.assert_array_equal(merged["weight"], want["edges"]["weight"])


@pytest.fixture()
def torch_stream_engine(engine):
    """the stages and the torch ops between them must share one stream (a real one: handle 0 means "engine's own")"""
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        engine.set_stream(stream.cuda_stream)
        yield engine
        torch.cuda.synchronize()
        engine.set_stream(None)


def _stages(eng, mode):
    return eng.a2a_stages() if mode == "alltoall" else eng.dist_stages()


@pytest.mark.parametrize("mode", ["alltoall", "allreduce"])
@pytest.mark.parametrize("world,n_ref,w", [(2, 1, 250), (3, 2, 100), (8, 1, 1000), (5, 3, 500)])
def test_lockstep_device_stages(torch_stream_engine, oracle, world, n_ref, w, mode):
    eng = torch_stream_engine
    dev = torch.device("cuda", 0)
    asms = _case(n_ref, 3_000_000)
    weights = [2.0] * n_ref + [1.0]
    full = [eng.sketch_buffers(s, o, 32, w) for s, o in asms]
    want = oracle.filter_and_edges([s.out_hash for s in full], [s.contig for s in full], weights)
    assert len(want["edges"]) > 1000
    hashes, contigs, keep = _rank_inputs(eng, asms, world, 32, w, dev)
    # sharded sketches concatenate to the single-GPU sketch
    for a in range(len(asms)):
        cat = torch.cat([hashes[r][a] for r in range(world)]).cpu().numpy().view(np.uint64)
        np.testing.assert_array_equal(cat, full[a].out_hash)
    shards = run_lockstep([_stages(eng, mode) for _ in range(world)], hashes, contigs, weights, dev)
    merged = merge_shards([s.fetch() for s in shards])
    _check(merged, want)
    single = eng.filter_and_edges(full, weights)
    np.testing.assert_array_equal(merged["edge_u"], single.edge_u)
    np.testing.assert_array_equal(merged["weight"], single.weight)
    assert sum(s.counts()[2] for s in shards) == len(single.edge_u)
    for s in shards:
        s.close()


@pytest.mark.parametrize("mode", ["alltoall", "allreduce"])
def test_lockstep_nothing_shared(torch_stream_engine, oracle, mode):
    eng = torch_stream_engine
    dev = torch.device("cuda", 0)
    a = synth.make_reference(400_000, n_chrom=3, seed=5)
    b = synth.make_reference(400_000, n_chrom=2, seed=6)
    asms = [(a[0], a[1]), (b[0], b[1])]
    hashes, contigs, keep = _rank_inputs(eng, asms, 4, 32, 100, dev)
    shards = run_lockstep([_stages(eng, mode) for _ in range(4)], hashes, contigs, [1.0, 1.0], dev)
    merged = merge_shards([s.fetch() for s in shards])
    assert len(merged["vertices"]) == 0 and len(merged["edge_u"]) == 0
    assert not any(k.any() for k in merged["keep"])


def _nccl_worker(rank, world, port, q, mode):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import ntjoin_b200
    import oracle_lib
    from ntjoin_b200.dist import TorchComm, distributed_filter_and_edges
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    eng = ntjoin_b200.Engine(rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    asms = _case(1, 4_000_000)
    weights = [2.0, 1.0]
    hashes, contigs, keep = _rank_inputs(eng, asms, world, 32, 500, dev)
    shard = distributed_filter_and_edges(_stages(eng, mode), hashes[rank], contigs[rank], weights, TorchComm(dev))
    gathered = [None] * world
    dist.all_gather_object(gathered, shard.fetch())
    ok = True
    if rank == 0:
        orc = oracle_lib.Oracle()
        full = [orc.sketch(s, o, 32, 500) for s, o in asms]
        want = orc.filter_and_edges([f["out_hash"] for f in full], [f["contig"] for f in full], weights)
        try:
            _check(merge_shards(gathered), want)
        except AssertionError as exc:
            print(exc)
            ok = False
    q.put((rank, ok))
    shard.close()
    dist.barrier()
    dist.destroy_process_group()
    eng.close()


def _run_nccl(world, mode):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 32500 + os.getpid() % 2000 + 17 * world
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port + (11 if mode == "alltoall" else 0), q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert sorted(r for r, _ in out) == list(range(world)) and all(ok for _, ok in out)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("mode", ["alltoall", "allreduce"])
def test_two_process_nccl(mode):
    _run_nccl(2, mode)


@pytest.mark.skipif(torch.cuda.device_count() < 3, reason="needs more than two GPUs")
@pytest.mark.parametrize("mode", ["alltoall", "allreduce"])
def test_all_devices_nccl(mode):
    """one rank per visible GPU (4 or 8 on a multi-GPU box), real NCCL collectives, merged shards against the oracle"""
    _run_nccl(torch.cuda.device_count(), mode)
