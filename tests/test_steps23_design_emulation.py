"""CPU: the single-GPU formulation of steps 2-3 (csrc/p2p.cu) restated stage by stage in numpy and checked
against the oracle on random and on sequence-derived inputs.  What this pins on the CPU is the DESIGN, not the kernels
(their parity tests need a GPU): vertex ids from the ascending kept hashes, one successor entry per (vertex, assembly),
"assembly b supports {v, x} iff succ_b[v] = x or succ_b[x] = v", ownership by the first supporting assembly, and the
PLACEMENT of the edges in the reference's formatted_edges order (bin/ntjoin_utils.py:115) without a sort -- the first edge
of a source is the one of the lowest assembly in the source's mask and carries the size of the source's block."""
import numpy as np
import pytest

from ntjoin_b200 import synth


def steps23_by_placement(hashes, contigs, weights):
    n_asm = len(hashes)
    key = np.concatenate(hashes).astype(np.uint64)
    asm = np.concatenate([np.full(len(h), a, dtype=np.int64) for a, h in enumerate(hashes)])
    ctg = np.concatenate(contigs).astype(np.int64)
    L = len(key)
    # stage "buckets": uniqueness per assembly, found exactly once in every assembly, vertex ranks (ascending hash)
    order = np.lexsort((asm, key))
    ks, as_ = key[order], asm[order]
    same_prev = np.zeros(L, dtype=bool)
    same_prev[1:] = (ks[1:] == ks[:-1]) & (as_[1:] == as_[:-1])
    same_next = np.zeros(L, dtype=bool)
    same_next[:-1] = same_prev[1:]
    uniq = np.zeros(L, dtype=bool)
    uniq[order] = ~(same_prev | same_next)
    run_start = np.ones(L, dtype=bool)
    run_start[1:] = ks[1:] != ks[:-1]
    run_id = np.cumsum(run_start) - 1
    run_len = np.bincount(run_id)
    pos_in_run = np.arange(L) - np.flatnonzero(run_start)[run_id]
    ok = (run_len[run_id] == n_asm) & (as_ == pos_in_run)                 # assemblies 0 .. n_asm-1, one element each
    good_run = np.bincount(run_id, weights=ok.astype(np.int64)) == n_asm
    keep = np.zeros(L, dtype=bool)
    keep[order] = good_run[run_id]
    vertices = np.unique(key[keep])
    # stage "adjacency": ordered survivors, sightings, successor table [vertex][assembly]
    cloc = np.flatnonzero(keep)
    cvid = np.searchsorted(vertices, key[cloc])
    a_j, c_j = asm[cloc], ctg[cloc]
    n_keep = len(cloc)
    eflag = np.zeros(n_keep, dtype=bool)
    eflag[:-1] = (a_j[1:] == a_j[:-1]) & (c_j[1:] == c_j[:-1])
    x_j = np.zeros(n_keep, dtype=np.int64)
    x_j[:-1] = cvid[1:]
    tab = np.zeros((len(vertices), n_asm), dtype=np.int64)
    tab[cvid, a_j] = np.where(eflag, x_j + 1, 0)                          # every (vertex, assembly) entry is written
    assert n_keep == len(vertices) * n_asm
    # stage "edges": support masks and ownership (first supporting assembly owns the edge, first-seen orientation)
    j = np.flatnonzero(eflag)
    v, x, a = cvid[j], x_j[j], a_j[j]
    mask = np.zeros(len(j), dtype=np.int64)
    for b in range(n_asm):
        mask |= ((tab[v, b] == x + 1) | (tab[x, b] == v + 1)).astype(np.int64) << b
    lowest = (mask & -mask).astype(np.int64)
    own = lowest == (1 << a)
    jo, vo, xo, ao, mo = j[own], v[own], x[own], a[own], mask[own]      # owned edges, in creation order
    smask = np.zeros(len(vertices), dtype=np.int64)
    np.bitwise_or.at(smask, vo, 1 << ao)
    # stage "finish": placement -- no sort
    sm = smask[vo]
    first = (sm & -sm) == (1 << ao)
    popc = np.array([bin(int(m)).count("1") for m in sm], dtype=np.int64)
    fcount = np.where(first, popc, 0)
    fprefix = np.concatenate([[0], np.cumsum(fcount)])[:-1]
    vstart = np.full(len(vertices), -1, dtype=np.int64)
    vstart[vo[first]] = fprefix[first]
    below = np.array([bin(int(m) & ((1 << int(b)) - 1)).count("1") for m, b in zip(sm, ao)], dtype=np.int64)
    o = np.where(first, fprefix, vstart[vo] + below)
    assert sorted(o.tolist()) == list(range(len(o)))                      # a permutation: every edge has its own place
    eu = np.empty(len(o), dtype=np.uint64)
    ev = np.empty(len(o), dtype=np.uint64)
    em = np.empty(len(o), dtype=np.uint32)
    ew = np.empty(len(o), dtype=np.float64)
    eu[o], ev[o], em[o] = vertices[vo], vertices[xo], mo
    ew[o] = [sum(weights[b] for b in range(n_asm) if m >> b & 1) for m in mo.tolist()]   # Python's sum(): assembly order, from int 0
    bounds = np.cumsum([0] + [len(h) for h in hashes])
    return {"uniq": [uniq[s:e] for s, e in zip(bounds[:-1], bounds[1:])], "keep": [keep[s:e] for s, e in zip(bounds[:-1], bounds[1:])],
            "vertices": vertices, "u": eu, "v": ev, "mask": em, "weight": ew}


def _compare(oracle, hashes, contigs, weights):
    want = oracle.filter_and_edges(hashes, contigs, weights)
    got = steps23_by_placement(hashes, contigs, weights)
    for a in range(len(hashes)):
        np.testing.assert_array_equal(got["uniq"][a], want["uniq"][a])
        np.testing.assert_array_equal(got["keep"][a], want["keep"][a])
    np.testing.assert_array_equal(got["vertices"], want["vertices"])
    np.testing.assert_array_equal(got["u"], want["edges"]["u"])
    np.testing.assert_array_equal(got["v"], want["edges"]["v"])
    np.testing.assert_array_equal(got["mask"], want["edges"]["support_mask"])
    np.testing.assert_array_equal(got["weight"], want["edges"]["weight"])
    return len(got["u"])


@pytest.mark.parametrize("n_asm,seed", [(1, 0), (2, 1), (3, 2), (4, 3), (5, 4)])
def test_placement_on_random_lists(oracle, n_asm, seed):
    rng = np.random.default_rng(300 + seed)
    edges = 0
    for _rep in range(15):
        pool = np.unique(rng.integers(1, 2**62, size=int(rng.integers(30, 300)), dtype=np.int64).astype(np.uint64))
        base = [pool[rng.permutation(len(pool))[:int(rng.integers(2, 25))]] for _ in range(int(rng.integers(2, 7)))]
        hashes, contigs = [], []
        for _a in range(n_asm):
            recs = []
            for b in base:
                r = b[::-1] if rng.random() < 0.5 else b
                if rng.random() < 0.5 and len(r) > 3:
                    c1, c2 = sorted(rng.integers(1, len(r), size=2))
                    pieces = [p for p in (r[:c1], r[c1:c2], r[c2:]) if len(p)]
                    r = np.concatenate([pieces[i][::-1] if rng.random() < 0.5 else pieces[i] for i in rng.permutation(len(pieces))])
                recs.append(r)
            if rng.random() < 0.5:
                recs.append(pool[rng.integers(0, len(pool), size=3)])
            recs = [recs[i] for i in rng.permutation(len(recs))]
            hashes.append(np.concatenate(recs))
            contigs.append(np.concatenate([np.full(len(r), c, dtype=np.uint32) for c, r in enumerate(recs)]))
        edges += _compare(oracle, hashes, contigs, [2.0, 1.0, 1.5, 0.1, 3.0][:n_asm])
    assert edges > 30


def test_placement_on_sequence(oracle):
    ref = synth.make_reference(3_000_000, n_chrom=5, seed=11, dup_frac=0.03, n_frac=0.004)
    other = ref[0].copy()
    rng = np.random.default_rng(5)
    idx = rng.integers(0, len(other), size=len(other) // 400)
    other[idx] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=len(idx))]
    tgt = synth.derive_target(ref[0], ref[1], min_len=3_000, max_len=150_000)
    sks = [oracle.sketch(s, o, 32, 100) for s, o in ((ref[0], ref[1]), (other, ref[1]), (tgt[0], tgt[1]))]
    n = _compare(oracle, [s["out_hash"] for s in sks], [s["contig"] for s in sks], [2.0, 2.0, 1.0])
    assert n > 10_000
