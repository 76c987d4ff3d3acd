"""Synthetic assemblies for the parity tests and bench.py.

All generators are deterministic in their seed (numpy PCG64) and return
(seq: uint8 ndarray of ASCII bases, offsets: uint64 ndarray of n_contigs+1, names: list[str]).
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP[_a] = _b

GRCH38_MBP = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]


def random_bases(n, rng):
    """n i.i.d. uniform ACGT bases."""
    out = np.empty(n, dtype=np.uint8)
    step = 1 << 26
    for s in range(0, n, step):
        e = min(n, s + step)
        out[s:e] = ACGT[rng.integers(0, 4, size=e - s, dtype=np.uint8)]
    return out


def make_reference(total_bp, n_chrom=10, seed=20251017, proportions=None, dup_frac=0.0, n_frac=0.0):
    """Reference assembly: n_chrom records of i.i.d. bases; optional duplicated 5 kb segments
    (copied 2-10x, exercises the uniqueness filter) and N runs of 100-50,000 bp (exercises A.4)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if proportions is None:
        proportions = [1.0] * n_chrom
    p = np.asarray(proportions[:n_chrom], dtype=np.float64)
    lens = np.maximum(1, (p / p.sum() * total_bp).astype(np.int64))
    offsets = np.zeros(n_chrom + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    n = int(offsets[-1])
    seq = random_bases(n, rng)
    if dup_frac > 0:
        seg = 5000
        n_src = max(1, int(n * dup_frac / seg / 6))
        for _ in range(n_src):
            s = int(rng.integers(0, max(1, n - seg)))
            for _c in range(int(rng.integers(1, 10))):
                d = int(rng.integers(0, max(1, n - seg)))
                seq[d:d + seg] = seq[s:s + seg]
    if n_frac > 0:
        target = int(n * n_frac)
        done = 0
        while done < target:
            ln = int(min(target - done + 100, np.exp(rng.uniform(np.log(100), np.log(50000)))))
            s = int(rng.integers(0, max(1, n - ln)))
            seq[s:s + ln] = ord("N")
            done += ln
    names = [f"chr{i + 1}" for i in range(n_chrom)]
    return seq, offsets, names


def revcomp(a):
    return _COMP[a[::-1]]


def derive_target(ref_seq, ref_offsets, seed=20251018, min_len=20_000, max_len=2_000_000, sub_rate=0.001, rc_frac=0.5):
    """Target assembly derived from a reference: cut into contigs with log-uniform lengths,
    rc_frac of them reverse-complemented, sub_rate substitutions, shuffled; names ctgNNNNNN."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pieces = []
    for c in range(len(ref_offsets) - 1):
        s, e = int(ref_offsets[c]), int(ref_offsets[c + 1])
        at = s
        while at < e:
            ln = int(np.exp(rng.uniform(np.log(min_len), np.log(max_len))))
            ln = min(ln, e - at)
            pieces.append((at, at + ln))
            at += ln
    order = rng.permutation(len(pieces))
    total = sum(b - a for a, b in pieces)
    out = np.empty(total, dtype=np.uint8)
    offsets = np.zeros(len(pieces) + 1, dtype=np.uint64)
    at = 0
    for i, pi in enumerate(order):
        a, b = pieces[pi]
        chunk = ref_seq[a:b]
        if rng.random() < rc_frac:
            chunk = revcomp(chunk)
        out[at:at + (b - a)] = chunk
        at += b - a
        offsets[i + 1] = at
    n_sub = int(total * sub_rate)
    if n_sub:
        idx = rng.integers(0, total, size=n_sub)
        keep = out[idx] != ord("N")
        out[idx[keep]] = ACGT[rng.integers(0, 4, size=int(keep.sum()), dtype=np.uint8)]
    names = [f"ctg{i:06d}" for i in range(len(pieces))]
    return out, offsets, names


def write_fasta(path, seq, offsets, names, width=0):
    with open(path, "wb") as f:
        for c, nm in enumerate(names):
            f.write(b">" + nm.encode() + b"\n")
            s = seq[int(offsets[c]):int(offsets[c + 1])].tobytes()
            if width:
                for i in range(0, len(s), width):
                    f.write(s[i:i + width] + b"\n")
            else:
                f.write(s + b"\n")
