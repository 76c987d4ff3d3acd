"""GPU parity of the peer-store formulation of steps 2-3 (csrc/p2p.cu, include/mxe.h: mxe_p2p_*).

world = 1 is what Engine.filter_and_edges runs by default (tests/test_gpu_filter.py, test_gpu_fullsize.py compare it
with the oracle and the reference-generated goldens).  Here: several ranks in ONE process on one device -- the same
kernels, the same symmetric workspaces and the same device-side barriers as the multi-process run, with the peers'
workspaces reached through plain pointers instead of CUDA IPC -- against the oracle and against world = 1, and real
processes (one per GPU, CUDA IPC over NVLink) when the box has more than one GPU."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle_lib
from ntjoin_b200 import synth
from ntjoin_b200.dist import merge_shards, shard_ranges

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(n_ref, total, seed=3):
    ref = synth.make_reference(total, n_chrom=6, seed=seed, dup_frac=0.03, n_frac=0.004)
    asms = [(ref[0], ref[1])]
    for r in range(1, n_ref):
        other = ref[0].copy()
        rng = np.random.default_rng(100 + r)
        idx = rng.integers(0, len(other), size=len(other) // 500)
        other[idx] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=len(idx))]
        asms.append((other, ref[1]))
    tgt = synth.derive_target(ref[0], ref[1], min_len=2_000, max_len=120_000)
    asms.append((tgt[0], tgt[1]))
    return asms


def _check(merged, want):
    np.testing.assert_array_equal(merged["vertices"], want["vertices"])
    for a in range(len(want["uniq"])):
        np.testing.assert_array_equal(np.asarray(merged["uniq"][a]).astype(bool), want["uniq"][a])
        np.testing.assert_array_equal(np.asarray(merged["keep"][a]).astype(bool), want["keep"][a])
    np.testing.assert_array_equal(merged["edge_u"], want["edges"]["u"])
    np.testing.assert_array_equal(merged["edge_v"], want["edges"]["v"])
    np.testing.assert_array_equal(merged["support"], want["edges"]["support_mask"])
    np.testing.assert_array_equal(merged["weight"], want["edges"]["weight"])


def _shard_sketches(eng, asms, world, k, w):
    """per rank: the sketches of its contiguous record range of every assembly"""
    ranges = shard_ranges([o for _, o in asms], world)
    out = []
    for r in range(world):
        sks = []
        for (seq, offs), (c0, c1) in zip(asms, ranges[r]):
            lo, hi = int(offs[c0]), int(offs[c1])
            sks.append(eng.sketch_buffers(seq[lo:hi].copy(), (offs[c0:c1 + 1] - offs[c0]).astype(np.uint64), k, w))
        out.append(sks)
    return out


def _fetch(res, world):
    d = dict(res.fetch())
    n_e = res.counts()[2]
    d["edge_key"] = res.edge_keys if (world > 1 and n_e) else np.arange(n_e, dtype=np.uint64)
    return d


def _lockstep(oracle, asms, weights, world, k, w, reps=2):
    """`world` ranks in this process, stage by stage; the merged shards must equal the oracle's steps 2-3"""
    import ntjoin_b200
    engines = [ntjoin_b200.Engine(0) for _ in range(world)]          # one engine (own stream, own arena) per simulated rank
    try:
        per_rank = [_shard_sketches(engines[r], asms, world, k, w)[r] for r in range(world)]
        total = sum(sk.n for sks in per_rank for sk in sks)
        groups = [engines[r].p2p(r, world, int(total * 1.2) + 1000, n_asm_max=len(asms)) for r in range(world)]
        bases = [g.workspace() for g in groups]
        for g in groups:
            g.connect_pointers(bases)
        full = [oracle.sketch(s, o, k, w) for s, o in asms]
        want = oracle.filter_and_edges([f["out_hash"] for f in full], [f["contig"] for f in full], weights)
        for rep in range(reps):                                       # the workspaces are reused across calls
            for r in range(world):
                groups[r].scatter_sketches(per_rank[r], weights)
            for stage in ("buckets", "adjacency", "edges"):
                for r in range(world):
                    getattr(groups[r], stage)()
            shards = [groups[r].finish() for r in range(world)]
            merged = merge_shards([_fetch(s, world) for s in shards])
            _check(merged, want)
            for s in shards:
                s.close()
        for g in groups:
            g.close()
        return merged
    finally:
        for e in engines:
            e.close()


@pytest.mark.parametrize("world,n_ref,w,records", [(1, 1, 100, 1), (2, 1, 100, 1), (3, 2, 250, 1), (5, 1, 50, 1), (8, 3, 100, 1),
                                                   (1, 3, 100, 1), (2, 1, 100, 0), (4, 2, 100, 0)])
def test_lockstep_ranks_one_device(oracle, world, n_ref, w, records, monkeypatch):
    """records = 0: the variant that keeps the successor tables at the vertex owners and reads / writes them with
    fine-grained peer accesses (MXE_P2P_RECORDS=0, read when the rank object is created)"""
    monkeypatch.setenv("MXE_P2P_RECORDS", str(records))
    merged = _lockstep(oracle, _case(n_ref, 1_500_000), [2.0] * n_ref + [1.0], world, 32, w)
    assert len(merged["vertices"]) > 1000 and len(merged["edge_u"]) > 1000


@pytest.mark.parametrize("world,bkmax", [(2, 512), (3, 1024)])
def test_repeated_sequence_several_ranks(oracle, world, bkmax, monkeypatch):
    """Half of every record is one 5 kb unit repeated 150 times: ~100 hashes with 900 copies per assembly, each far more
    than a bucket CTA holds in shared memory.  The owner places buckets at exact offsets and the bucket kernel keeps two
    copies of every (hash, assembly) pair while loading, so the multi-rank path answers (world = 1 has the sort-based
    fallback, test_bucket_overflow_falls_back): flags, vertices and edges equal the oracle's."""
    monkeypatch.setenv("MXE_P2P_BKMAX", str(bkmax))
    rng = np.random.default_rng(7)
    unit = synth.random_bases(5000, rng)
    parts, offs = [], [0]
    for _c in range(6):
        parts += [synth.random_bases(400_000, rng)] + [unit] * 150 + [synth.random_bases(350_000, rng)]
        offs.append(offs[-1] + 750_000 + 150 * 5000)
    ref = (np.concatenate(parts), np.array(offs, dtype=np.uint64))
    tgt = synth.derive_target(ref[0], ref[1], min_len=20_000, max_len=400_000)
    merged = _lockstep(oracle, [ref, (tgt[0], tgt[1])], [2.0, 1.0], world, 32, 100)
    assert len(merged["vertices"]) > 10_000
    n_dup = int((~np.asarray(merged["uniq"][0]).astype(bool)).sum())
    assert n_dup > 50_000                                             # the repeats really are there


def test_bucket_overflow_falls_back(engine, oracle):
    """one hash repeated thousands of times overflows its bucket: mxe_filter_and_edges must still answer (sort-based path)"""
    rng = np.random.default_rng(1)
    unit = synth.random_bases(5000, rng)
    seq = np.concatenate([unit] * 3000 + [synth.random_bases(200_000, rng)])
    offs = np.array([0, len(seq)], dtype=np.uint64)
    ref = oracle.sketch(seq, offs, 32, 100)
    sk = engine.sketch_buffers(seq, offs, 32, 100)
    res = engine.filter_and_edges([sk, sk], [1.0, 1.0])
    want = oracle.filter_and_edges([ref["out_hash"]] * 2, [ref["contig"]] * 2, [1.0, 1.0])
    np.testing.assert_array_equal(res.vertices, want["vertices"])
    np.testing.assert_array_equal(res.edge_u, want["edges"]["u"])
    np.testing.assert_array_equal(res.uniq[0].astype(bool), want["uniq"][0])


def _ipc_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import ntjoin_b200
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)     # host plumbing only: 64-byte handles and the verdict
    eng = ntjoin_b200.Engine(rank)
    asms = _case(1, 4_000_000)
    weights = [2.0, 1.0]
    sks = _shard_sketches(eng, asms, world, 32, 500)[rank]
    tot = torch.tensor([sum(sk.n for sk in sks)], dtype=torch.int64)
    dist.all_reduce(tot)
    grp = eng.p2p(rank, world, int(tot.item() * 1.2) + 1000, n_asm_max=2)
    handles = [None] * world
    dist.all_gather_object(handles, grp.handle())
    grp.connect(handles)
    ok = True
    for rep in range(2):
        shard = grp.run(sks, weights)
        gathered = [None] * world
        dist.all_gather_object(gathered, _fetch(shard, world))
        if rank == 0:
            orc = oracle_lib.Oracle()
            full = [orc.sketch(s, o, 32, 500) for s, o in asms]
            want = orc.filter_and_edges([f["out_hash"] for f in full], [f["contig"] for f in full], weights)
            try:
                _check(merge_shards(gathered), want)
            except AssertionError as exc:
                print(exc)
                ok = False
        shard.close()
    q.put((rank, ok))
    dist.barrier()
    grp.close()
    dist.destroy_process_group()
    eng.close()


def _run_ipc(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 34500 + os.getpid() % 2000 + 13 * world
    procs = [ctx.Process(target=_ipc_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert sorted(r for r, _ in out) == list(range(world)) and all(ok for _, ok in out)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_processes_ipc():
    _run_ipc(2)


@pytest.mark.skipif(torch.cuda.device_count() < 3, reason="needs more than two GPUs")
def test_all_devices_ipc():
    _run_ipc(torch.cuda.device_count())
