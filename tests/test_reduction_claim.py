"""CPU: the claim behind the oversized-bucket path of csrc/p2p.cu ("Repeated sequence") checked with
the oracle alone: dropping all but two copies of every (hash, assembly) pair changes nothing -- the dropped copies are
"not unique, not kept" either way, and flags, vertices and edges of everything else stay what they were
(bin/ntjoin_utils.py:182-187 only asks "once" or "more than once")."""
import numpy as np
import pytest


@pytest.mark.parametrize("n_asm,seed", [(1, 0), (2, 1), (3, 2), (4, 3)])
def test_two_copies_per_hash_and_assembly_are_enough(oracle, n_asm, seed):
    rng = np.random.default_rng(50 + seed)
    pool = np.unique(rng.integers(1, 2**62, size=3000, dtype=np.int64).astype(np.uint64))
    heavy = pool[:12]
    hashes, contigs = [], []
    for _a in range(n_asm):
        recs = []
        for _c in range(int(rng.integers(3, 9))):
            body = pool[rng.permutation(len(pool))[:int(rng.integers(50, 400))]]
            ins = rng.integers(0, len(body), size=int(rng.integers(0, 300)))          # heavy hitters sprinkled in
            body = np.insert(body, ins, heavy[rng.integers(0, len(heavy), size=len(ins))])
            recs.append(body)
        hashes.append(np.concatenate(recs))
        contigs.append(np.concatenate([np.full(len(r), c, dtype=np.uint32) for c, r in enumerate(recs)]))
    weights = [2.0, 1.0, 1.5, 0.5][:n_asm]
    full = oracle.filter_and_edges(hashes, contigs, weights)
    # keep the first two copies of every (hash, assembly) pair, in any order the bucket kernel might see them
    red_h, red_c, kept_idx = [], [], []
    for h, c in zip(hashes, contigs):
        order = rng.permutation(len(h))
        seen = {}
        keep = np.zeros(len(h), dtype=bool)
        for i in order.tolist():
            n = seen.get(int(h[i]), 0)
            if n < 2:
                keep[i] = True
                seen[int(h[i])] = n + 1
        kept_idx.append(np.flatnonzero(keep))
        red_h.append(h[keep])
        red_c.append(c[keep])
    assert sum(len(h) - len(r) for h, r in zip(hashes, red_h)) > 100           # something was dropped
    red = oracle.filter_and_edges(red_h, red_c, weights)
    for a in range(n_asm):
        u = np.zeros(len(hashes[a]), dtype=bool)
        k = np.zeros(len(hashes[a]), dtype=bool)
        u[kept_idx[a]] = red["uniq"][a]
        k[kept_idx[a]] = red["keep"][a]                                          # dropped copies: mark 0
        np.testing.assert_array_equal(u, full["uniq"][a])
        np.testing.assert_array_equal(k, full["keep"][a])
    np.testing.assert_array_equal(red["vertices"], full["vertices"])
    np.testing.assert_array_equal(red["edges"], full["edges"])
