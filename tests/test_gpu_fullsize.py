"""GPU parity at BASELINE.json's full sizes.
configs[1] (100 Mbp + 100 Mbp): direct comparison with the oracle.
configs[2] (3 Gbp): size-independent properties -- order, idempotence, threshold independence, density, and
slice consistency (minimizers further than w+k valid k-mers from a cut are a local property, so the oracle run
on slices must reproduce the interior of the full-size result exactly)."""
import os

import numpy as np
import pytest

from ntjoin_b200 import synth

pytestmark = pytest.mark.gpu
K, W = 32, 1000


def test_config2_full(engine, oracle):
    rseq, roffs, rn = synth.make_reference(100_000_000, n_chrom=10)
    tseq, toffs, tn = synth.derive_target(rseq, roffs)
    cores = os.cpu_count() or 1
    sks = []
    for seq, offs, names in ((rseq, roffs, rn), (tseq, toffs, tn)):
        ref = oracle.sketch(seq, offs, K, W, threads=cores)
        sk = engine.sketch_buffers(seq, offs, K, W, names=names)
        assert sk.n == len(ref) and abs(sk.n - 2e8 / (W + 1)) < 0.02 * 2e8 / (W + 1)
        np.testing.assert_array_equal(sk.out_hash, ref["out_hash"])
        np.testing.assert_array_equal(sk.pos, ref["pos"].astype(np.uint32))
        np.testing.assert_array_equal(sk.contig, ref["contig"])
        np.testing.assert_array_equal(sk.forward.astype(np.uint32), ref["forward"])
        sks.append(sk)
    res = engine.filter_and_edges(sks, [2.0, 1.0])
    want = oracle.filter_and_edges([s.out_hash for s in sks], [s.contig for s in sks], [2.0, 1.0])
    np.testing.assert_array_equal(res.vertices, want["vertices"])
    np.testing.assert_array_equal(res.edge_u, want["edges"]["u"])
    np.testing.assert_array_equal(res.edge_v, want["edges"]["v"])
    np.testing.assert_array_equal(res.support, want["edges"]["support_mask"])
    np.testing.assert_array_equal(res.weight, want["edges"]["weight"])
    assert len(res.vertices) > 150_000


def test_config3_properties(engine, oracle):
    import torch
    dev = torch.device("cuda", 0)
    n = 3_000_000_000
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    seq = torch.empty(n, dtype=torch.uint8, device=dev)
    for s in range(0, n, 1 << 28):
        e = min(n, s + (1 << 28))
        seq[s:e] = lut[torch.randint(0, 4, (e - s,), dtype=torch.uint8, device=dev, generator=g).long()]
    rng = np.random.default_rng(3)
    for _ in range(300):                                   # N runs incl. long ones, and duplicated segments
        s = int(rng.integers(0, n - 200_000)); ln = int(np.exp(rng.uniform(np.log(10), np.log(100_000))))
        seq[s:s + ln] = ord("N")
    for _ in range(2000):
        s, d = (int(x) for x in rng.integers(0, n - 5000, 2))
        seq[d:d + 5000] = seq[s:s + 5000].clone()
    lens = np.array(synth.GRCH38_MBP, dtype=np.float64)
    offs = np.concatenate([[0], np.cumsum((lens / lens.sum() * n).astype(np.int64))]).astype(np.uint64)
    offs[-1] = n
    torch.cuda.synchronize()

    sk = engine.sketch_device(seq.data_ptr(), offs, K, W)
    oh, pos, ctg = sk.out_hash.copy(), sk.pos.copy(), sk.contig.copy()
    c = sk.counts()
    assert c["bases"] == n and c["contigs"] == 24
    # density 2/(w+1) within 2 %
    assert abs(len(oh) - 2 * c["valid_kmers"] / (W + 1)) < 0.02 * 2 * n / (W + 1)
    # ordered by (record, pos), strictly increasing inside a record
    key = ctg.astype(np.uint64) << np.uint64(32) | pos.astype(np.uint64)
    assert np.all(key[1:] > key[:-1])
    # idempotent, and independent of the candidate threshold (performance knob only)
    engine.set_option("tau", 7.0)
    try:
        sk2 = engine.sketch_device(seq.data_ptr(), offs, K, W)
        np.testing.assert_array_equal(sk2.out_hash, oh)
        np.testing.assert_array_equal(sk2.pos, pos)
        np.testing.assert_array_equal(sk2.contig, ctg)
    finally:
        engine.set_option("tau", 9.0)
    # slice consistency against the oracle
    gpos = offs[ctg] + pos.astype(np.uint64)
    margin = 4 * (W + K)
    for rec in (0, 7, 23):
        for frac in (0.0, 0.37, 0.999):
            a0, b0 = int(offs[rec]), int(offs[rec + 1])
            a = a0 + int((b0 - a0 - 400_000) * frac)
            b = a + 400_000
            sl = seq[a:b].cpu().numpy()
            ref = oracle.sketch(sl, np.array([0, len(sl)], dtype=np.uint64), K, W)
            lo = a if a == a0 else a + margin
            hi = b if b == b0 else b - margin
            if (sl[:margin] == ord("N")).any() or (sl[-margin:] == ord("N")).any():
                continue                                   # an N run at the cut widens the needed margin
            mine = gpos[(gpos >= lo) & (gpos < hi)] - a
            theirs = ref["pos"][(ref["pos"] + a >= lo) & (ref["pos"] + a < hi)]
            np.testing.assert_array_equal(mine, theirs)
            sel = (gpos >= lo) & (gpos < hi)
            np.testing.assert_array_equal(oh[sel], ref["out_hash"][(ref["pos"] + a >= lo) & (ref["pos"] + a < hi)])


def test_config3_full_bit_exact(engine, oracle):
    """BASELINE configs[2] at FULL size, the north_star's acceptance sentence: both 3 Gbp assemblies of bench.py's
    workload in the with-N variant (reference with 0.5 % of its bases in N runs and 2 % in duplicated 5 kb segments;
    target = its contigs cut, half reverse-complemented, mutated, shuffled) are sketched on the GPU and by the oracle
    with every host thread, and ALL minimizer tuples must be equal; then steps 2-3 (6 M x 2 minimizers) against the
    oracle's: flags, vertices, and the weighted edge list in the reference's order."""
    import sys
    import types
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    dev = torch.device("cuda", 0)
    spec = bench.workload_spec(types.SimpleNamespace(workload="c3", bases=0))
    assemblies = bench.gen_assemblies_gpu(spec, True, dev)            # with_n=True
    cores = os.cpu_count() or 1
    sks, refs = [], []
    for seq, offs in assemblies:
        sk = engine.sketch_device(seq.data_ptr(), offs, K, W)
        host = seq.cpu().numpy()
        ref = oracle.sketch(host, offs, K, W, threads=cores)
        del host
        assert sk.n == len(ref) and sk.n > 5_000_000
        np.testing.assert_array_equal(sk.out_hash, ref["out_hash"])
        np.testing.assert_array_equal(sk.min_hash, ref["min_hash"])
        np.testing.assert_array_equal(sk.pos, ref["pos"].astype(np.uint32))
        np.testing.assert_array_equal(sk.contig, ref["contig"])
        np.testing.assert_array_equal(sk.forward.astype(np.uint32), ref["forward"])
        assert sk.counts()["gap_windows"] > 0                          # the N runs put the dense gap path to work
        sks.append(sk)
        refs.append(ref)
    del assemblies
    res = engine.filter_and_edges(sks, [2.0, 1.0])
    want = oracle.filter_and_edges([r["out_hash"] for r in refs], [r["contig"] for r in refs], [2.0, 1.0])
    for a in range(2):
        np.testing.assert_array_equal(res.uniq[a].astype(bool), want["uniq"][a])
        np.testing.assert_array_equal(res.keep[a].astype(bool), want["keep"][a])
    np.testing.assert_array_equal(res.vertices, want["vertices"])
    np.testing.assert_array_equal(res.edge_u, want["edges"]["u"])
    np.testing.assert_array_equal(res.edge_v, want["edges"]["v"])
    np.testing.assert_array_equal(res.support, want["edges"]["support_mask"])
    np.testing.assert_array_equal(res.weight, want["edges"]["weight"])
    assert len(res.vertices) > 5_000_000 and len(res.edge_u) > 5_000_000
    for sk in sks:
        sk.close()
    res.close()
