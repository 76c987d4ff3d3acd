"""CPU: the exact-hash kernels of the sketch tail (cand_hash_pos_kernel, final_eval_pos_kernel, gap_kernel) do not roll:
ntHash is XOR-linear in the bases, so a k-mer's forward and reverse hash is the XOR of one table entry per 4-base group,
each entry carrying the group's rotation already (hash_pos_tables_kernel), plus single-base steps for the k % 4 trailing
bases.  The table construction and the look-up scheme of ntjoin_b200/csrc/sketch_kernels.cuh are restated here in Python
integers and compared with the oracle's rolling hashes (device base codes: A 0, C 1, T 2, G 3; complement = code ^ 2)."""
import numpy as np
import pytest

M33, M31 = (1 << 33) - 1, (1 << 31) - 1
SEED = [0x3c8bfbb395c60474, 0x3193c18562a02b4c, 0x295549f54be24456, 0x20323ed082572324]      # A C T G
CODE = {"A": 0, "C": 1, "T": 2, "G": 3}


def srol1(x):
    lo, hi = x & M33, x >> 33
    lo = ((lo << 1) | (lo >> 32)) & M33
    hi = ((hi << 1) | (hi >> 30)) & M31
    return (hi << 33) | lo


def sror1(x):
    lo, hi = x & M33, x >> 33
    lo = (lo >> 1) | ((lo & 1) << 32)
    hi = (hi >> 1) | ((hi & 1) << 30)
    return (hi << 33) | lo


def tables(k):
    """f4 / r4 as build_hash_tabs, PF / PR as hash_pos_tables_kernel, s1 as stage_pos_tables"""
    f4, r4 = [], []
    for v in range(256):
        f = r = 0
        for j in range(4):
            c = (v >> (2 * j)) & 3
            f = srol1(f) ^ SEED[c]
            sc = SEED[c ^ 2]
            for _ in range(j):
                sc = srol1(sc)
            r ^= sc
        for _ in range(k - 4):
            r = srol1(r)
        f4.append(f)
        r4.append(r)
    G = k // 4
    PF, PR = [], []
    for g in range(G):
        rowf, rowr = [], []
        for v in range(256):
            f, r = f4[v], r4[v]
            for _ in range(4 * (G - 1 - g)):
                f, r = srol1(f), sror1(r)
            rowf.append(f)
            rowr.append(r)
        PF.append(rowf)
        PR.append(rowr)
    s1 = list(SEED)
    for c in range(4):
        sc = SEED[c ^ 2]
        for _ in range(k - 1):
            sc = srol1(sc)
        s1.append(sc)
    return PF, PR, s1


def hash_by_tables(kmer, PF, PR, s1):
    k = len(kmer)
    codes = [CODE[ch] for ch in kmer]
    f = r = 0
    G = k // 4
    for g in range(G):
        v = sum(codes[4 * g + j] << (2 * j) for j in range(4))
        f ^= PF[g][v]
        r ^= PR[g][v]
    for c in codes[4 * G:]:
        f = srol1(f) ^ s1[c]
        r = sror1(r) ^ s1[4 + c]
    return f, r


@pytest.mark.parametrize("k", [4, 15, 21, 24, 32, 40, 43])
def test_position_specific_tables_give_the_rolling_hashes(oracle, k):
    rng = np.random.default_rng(k)
    PF, PR, s1 = tables(k)
    for _ in range(200):
        kmer = "".join("ACGT"[i] for i in rng.integers(0, 4, size=k))
        fwd, rev, _h0, _h1 = oracle.kmer_hashes(kmer)
        assert hash_by_tables(kmer, PF, PR, s1) == (fwd, rev), kmer
    assert srol1(sror1(0x123456789abcdef0)) == 0x123456789abcdef0
