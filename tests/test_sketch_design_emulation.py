"""CPU: the claim behind the sketch kernels, restated in numpy and checked against the oracle: the ordered
minimizers of a record can be found WITHOUT a sliding window over all k-mers -- take any superset of the k-mers whose
top 31 hash bits are at most T ("candidates"), select per window over the candidates only, accept a window whose
candidate arg-min really is below the threshold, and evaluate only the remaining windows ("gaps") densely.  The result
must not depend on T nor on which extra k-mers the superset contains.  (The kernels' own parity tests need a GPU.)"""
import numpy as np
import pytest

from ntjoin_b200 import synth


def minimizers_by_prefilter(oracle, seq, k, w, tau, rng, extra=0.02):
    """positions of the minimizers of ONE record, by the candidate / gap scheme"""
    n = len(seq)
    text = bytes(seq).upper()
    valid = np.zeros(n, dtype=bool)
    h0 = np.zeros(n, dtype=np.uint64)
    ok = np.frombuffer(text, dtype=np.uint8)
    good = np.isin(ok, np.frombuffer(b"ACGT", dtype=np.uint8))
    run = 0
    for p in range(n - 1, -1, -1):                      # valid k-mer start: k good bases from p on
        run = run + 1 if good[p] else 0
        valid[p] = run >= k
    for p in np.flatnonzero(valid):
        h0[p] = oracle.kmer_hashes(text[p:p + k])[2]
    vpos = np.flatnonzero(valid)                        # ordinal -> position
    if len(vpos) < w:
        return []
    hv = h0[vpos]
    T = int(tau * 2**31 / w)
    cand = (hv >> np.uint64(33)) <= np.uint64(T)
    cand |= rng.random(len(vpos)) < extra               # any superset will do
    cand_ord = np.flatnonzero(cand)
    out = set()
    for j in range(w - 1, len(vpos)):                   # window = ordinals j-w+1 .. j
        lo = j - w + 1
        c = cand_ord[(cand_ord >= lo) & (cand_ord <= j)]
        best = None
        if len(c):
            hc = hv[c]
            m = hc.min()
            best = int(c[np.flatnonzero(hc == m)[-1]])  # rightmost of the smallest
            if int(hv[best] >> np.uint64(33)) > T:
                best = None                             # the true minimum may be a k-mer that is no candidate
        if best is None:                                # gap: dense, exact
            hw_ = hv[lo:j + 1]
            m = hw_.min()
            best = lo + int(np.flatnonzero(hw_ == m)[-1])
        out.add(int(vpos[best]))
    return sorted(out)


@pytest.mark.parametrize("k,w", [(32, 50), (15, 10), (24, 100)])
def test_prefilter_is_exact_and_threshold_independent(oracle, k, w):
    rng = np.random.default_rng(9)
    ref = synth.make_reference(12_000, n_chrom=1, seed=77, dup_frac=0.05, n_frac=0.02)
    seq = ref[0].copy()
    seq[3000:3004] = np.frombuffer(b"acgt", dtype=np.uint8)          # lower case counts as its upper case
    seq[5000:5200] = ord("A")                                        # a low-complexity run: many equal hashes (ties)
    offs = np.array([0, len(seq)], dtype=np.uint64)
    want = oracle.sketch(seq, offs, k, w)["pos"].astype(np.int64).tolist()
    assert len(want) > 50
    for tau in (0.5, 3.0, 9.0, 40.0):
        got = minimizers_by_prefilter(oracle, seq, k, w, tau, rng)
        assert got == want, (k, w, tau)


@pytest.mark.parametrize("w", [250, 1000, 5000])
def test_twelve_bit_lane_test_is_a_superset(w):
    """scan_bs2_kernel decides from the top 12 bits of the two 31-bit lanes alone: with
    hash0 = fwd + rev (mod 2^64), t = hash0 >> 33 = (f31 + r31 + c) mod 2^31 where c is the carry out of the low 33 bits,
    and S = (f31 >> 19) + (r31 >> 19) + 1 (mod 2^12) satisfies  t <= T  =>  S <= (T >> 19) + 1: every k-mer below the
    threshold is flagged, whatever the dropped bits and carries are; the flagged set is only a few per cent larger."""
    rng = np.random.default_rng(w)
    n = 2_000_000
    fwd = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    rev = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    # values at the edges of the carry cases
    fwd[:4] = [0, 2**64 - 1, (1 << 33) - 1, 1 << 33]
    rev[:4] = [0, 1, 1, (1 << 64) - (1 << 33)]
    h0 = fwd + rev                                                       # wraps mod 2^64
    t = (h0 >> np.uint64(33)).astype(np.int64)
    f31, r31 = (fwd >> np.uint64(33)).astype(np.int64), (rev >> np.uint64(33)).astype(np.int64)
    c = (((fwd & np.uint64((1 << 33) - 1)).astype(np.int64) + (rev & np.uint64((1 << 33) - 1)).astype(np.int64)) >> 33)
    assert np.array_equal(t, (f31 + r31 + c) & ((1 << 31) - 1))          # the two rotation groups never mix
    T = int(9.0 * 2**31 / w)
    S = ((f31 >> 19) + (r31 >> 19) + 1) & 0xFFF
    Q = min((T >> 19) + 1, 0xFFF)
    flagged = S <= Q
    below = t <= T
    assert not np.any(below & ~flagged)
    assert below.sum() > 0 and flagged.sum() >= below.sum()
    if w >= 1000:
        assert flagged.sum() < 1.6 * below.sum() + 50
