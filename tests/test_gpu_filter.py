"""GPU parity for steps 2-3: uniqueness, intersection, edge list (C ABI) vs golden vectors made by the
reference's own Python (tests/golden/make_golden.py) and vs the CPU oracle on synthetic data."""
import glob
import json
import os

import numpy as np
import pytest

import oracle_lib
from ntjoin_b200 import synth

pytestmark = pytest.mark.gpu


def _cases(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "steps23_*.json")))


def test_golden_steps23(engine, golden_dir):
    files = _cases(golden_dir)
    assert files
    for path in files:
        g = json.load(open(path))
        sks = [engine.sketch_file(os.path.join(golden_dir, "inputs", f), g["k"], g["w"]) for f in g["files"]]
        res = engine.filter_and_edges(sks, g["weights"])
        for a, sk in enumerate(sks):
            # read_minimizers: mx_info keys = unique hashes, with (contig, pos)
            info = {str(h): [sk.names[c], int(p)] for h, c, p in zip(sk.out_hash[res.uniq[a]], sk.contig[res.uniq[a]], sk.pos[res.uniq[a]])}
            assert info == g["read_minimizers"][a]["mx_info"], path
            # filter_minimizers: per record ordered survivor lists (records without minimizers have no list)
            lists = []
            for c in range(len(sk.names)):
                sel = sk.contig == c
                if sel.any():
                    lists.append([str(h) for h in sk.out_hash[sel & res.keep[a]]])
            assert lists == g["filter_minimizers"][a], path
        assert [str(v) for v in res.vertices] == g["vertices"], path
        assert [[str(u), str(v)] for u, v in zip(res.edge_u, res.edge_v)] == g["edges"], path
        sup = [[a for a in range(len(sks)) if m >> a & 1] for m in res.support]
        assert sup == g["support"], path
        assert [float(x) for x in res.weight] == g["weight"], path


def test_synthetic_vs_oracle(engine, oracle):
    rseq, roffs, _ = synth.make_reference(6_000_000, n_chrom=6, dup_frac=0.02, n_frac=0.005)
    asms = [(rseq, roffs)]
    for s in (1, 2):
        asms.append(synth.derive_target(rseq, roffs, seed=100 + s, min_len=5000, max_len=400_000, sub_rate=0.002)[:2])
    weights = [2.0, 1.5, 1.0]
    sks = [engine.sketch_buffers(s, o, 32, 250) for s, o in asms]
    res = engine.filter_and_edges(sks, weights)
    want = oracle.filter_and_edges([s.out_hash for s in sks], [s.contig for s in sks], weights)
    for a in range(3):
        np.testing.assert_array_equal(res.uniq[a], want["uniq"][a])
        np.testing.assert_array_equal(res.keep[a], want["keep"][a])
    np.testing.assert_array_equal(res.vertices, want["vertices"])
    np.testing.assert_array_equal(res.edge_u, want["edges"]["u"])
    np.testing.assert_array_equal(res.edge_v, want["edges"]["v"])
    np.testing.assert_array_equal(res.support, want["edges"]["support_mask"])
    np.testing.assert_array_equal(res.weight, want["edges"]["weight"])
    assert len(res.vertices) > 10000 and len(res.edge_u) > 10000
    assert 7 in set(np.unique(res.support)) and len(set(np.unique(res.support))) >= 2


@pytest.mark.parametrize("n_top", [40, 40_000])
def test_hash_sort_tied_top_bits(engine, oracle, n_top):
    """steps 2-3 sort only the top 32 bits and repair tied runs in place; long mixed runs force the full 64-bit
    sort (n_top=40: runs of thousands of different hashes sharing their top 32 bits), short ones the fix-up path"""
    import torch
    rng = np.random.default_rng(n_top)
    tops = rng.integers(0, 2**32, n_top, dtype=np.uint64) << np.uint64(32)
    base = tops[rng.integers(0, n_top, 120_000)] | rng.integers(0, 2**12, 120_000, dtype=np.uint64)
    hashes, contigs = [], []
    for a in range(3):
        h = base.copy()
        rng.shuffle(h)
        h = h[: 100_000 + 1000 * a]
        hashes.append(h)
        contigs.append(np.sort(rng.integers(0, 50, len(h))).astype(np.uint32))
    weights = [1.0, 0.25, 3.0]
    dh = [torch.from_numpy(h.view(np.int64)).cuda() for h in hashes]
    dc = [torch.from_numpy(c.view(np.int32)).cuda() for c in contigs]
    torch.cuda.synchronize()
    res = engine.filter_and_edges_device([t.data_ptr() for t in dh], [t.data_ptr() for t in dc], [len(h) for h in hashes], weights)
    want = oracle.filter_and_edges(hashes, contigs, weights)
    for a in range(3):
        np.testing.assert_array_equal(res.uniq[a], want["uniq"][a])
        np.testing.assert_array_equal(res.keep[a], want["keep"][a])
    np.testing.assert_array_equal(res.vertices, want["vertices"])
    np.testing.assert_array_equal(res.edge_u, want["edges"]["u"])
    np.testing.assert_array_equal(res.edge_v, want["edges"]["v"])
    np.testing.assert_array_equal(res.support, want["edges"]["support_mask"])
    np.testing.assert_array_equal(res.weight, want["edges"]["weight"])
    assert len(res.vertices) > 100


def test_config4_four_way(engine, oracle):
    """BASELINE configs[3] scaled down: target + 3 references (0.1-0.5 % divergence), k=32 w=500, weights 2 2 2 / 1"""
    anc, aoffs, _ = synth.make_reference(8_000_000, n_chrom=6, dup_frac=0.02)
    asms = []
    for i, rate in enumerate((0.001, 0.003, 0.005)):
        asms.append(synth.derive_target(anc, aoffs, seed=500 + i, min_len=200_000, max_len=3_000_000, sub_rate=rate, rc_frac=0.2)[:2])
    asms.append(synth.derive_target(anc, aoffs, seed=600, min_len=10_000, max_len=500_000, sub_rate=0.002)[:2])      # target last
    weights = [2.0, 2.0, 2.0, 1.0]
    sks = [engine.sketch_buffers(s, o, 32, 500) for s, o in asms]
    res = engine.filter_and_edges(sks, weights)
    want = oracle.filter_and_edges([s.out_hash for s in sks], [s.contig for s in sks], weights)
    for a in range(4):
        np.testing.assert_array_equal(res.uniq[a], want["uniq"][a])
        np.testing.assert_array_equal(res.keep[a], want["keep"][a])
    np.testing.assert_array_equal(res.vertices, want["vertices"])
    np.testing.assert_array_equal(res.edge_u, want["edges"]["u"])
    np.testing.assert_array_equal(res.edge_v, want["edges"]["v"])
    np.testing.assert_array_equal(res.support, want["edges"]["support_mask"])
    np.testing.assert_array_equal(res.weight, want["edges"]["weight"])
    masks = set(int(m) for m in np.unique(res.support))
    assert 15 in masks and len(masks) >= 4 and set(np.unique(res.weight)) >= {7.0}
    # n=2 filtering downstream (bin/ntjoin.py:80-89) keeps edges with weight >= 2: every edge seen only in the target is dropped
    assert (res.weight[res.support == 8] == 1.0).all()


@pytest.mark.parametrize("bits", [24, 32, 40])
def test_sort_width_independence(engine, oracle, bits):
    """the radix sort covers the top 24/32/40 hash bits and the fix-up the rest: same result for every split"""
    rseq, roffs, _ = synth.make_reference(2_000_000, n_chrom=4, dup_frac=0.05)
    asms = [(rseq, roffs), synth.derive_target(rseq, roffs, seed=5, min_len=4000, max_len=100_000)[:2]]
    sks = [engine.sketch_buffers(s, o, 32, 50) for s, o in asms]
    want = oracle.filter_and_edges([s.out_hash for s in sks], [s.contig for s in sks], [2.0, 1.0])
    engine.set_option("sort_bits", bits)
    try:
        res = engine.filter_and_edges(sks, [2.0, 1.0])
        np.testing.assert_array_equal(res.vertices, want["vertices"])
        np.testing.assert_array_equal(res.edge_u, want["edges"]["u"])
        np.testing.assert_array_equal(res.edge_v, want["edges"]["v"])
        np.testing.assert_array_equal(res.support, want["edges"]["support_mask"])
        for a in range(2):
            np.testing.assert_array_equal(res.uniq[a], want["uniq"][a])
    finally:
        engine.set_option("sort_bits", 0)
