"""GPU parity: CUDA sketch (through the C ABI) vs the CPU oracle and the reference's golden TSVs."""
import os

import numpy as np
import pytest

import oracle_lib
from ntjoin_b200 import synth

pytestmark = pytest.mark.gpu

FIXTURES = ["ref.fa", "ref.multiple.fa", "scaf.f-f.fa", "scaf.f-f.termN.unassigned.fa", "scaf.multiple.fa",
            "scaf.more_seqs.fa", "scaf.f-f.overlapping.fa", "scaf.r-r.fa"]
DEFAULT_VARIANT = 4      # pack2 + bit-sliced scan (falls back to cand31 / generic where it does not apply)
KW = [(32, 1000), (32, 500), (32, 250), (15, 10), (24, 100), (40, 50), (21, 33), (32, 5000), (4, 3)]


def assert_same(sk, ref):
    assert sk.n == len(ref), f"minimizer count {sk.n} != oracle {len(ref)}"
    np.testing.assert_array_equal(sk.contig, ref["contig"])
    np.testing.assert_array_equal(sk.pos.astype(np.uint64), ref["pos"])
    np.testing.assert_array_equal(sk.min_hash, ref["min_hash"])
    np.testing.assert_array_equal(sk.out_hash, ref["out_hash"])
    np.testing.assert_array_equal(sk.forward.astype(np.uint32), ref["forward"])


@pytest.mark.parametrize("fname", ["ref.fa", "scaf.f-f.fa"])
def test_golden_tsv_bytes(engine, golden_dir, tmp_path, fname):
    """tests/expected_outputs/*.k32.w1000.tsv of the reference: byte-equal under canonical=min, --pos."""
    sk = engine.sketch_file(os.path.join(golden_dir, "inputs", fname), 32, 1000, canonical="min")
    out = tmp_path / "o.tsv"
    sk.write_tsv(out, pos=True, strand=False, seq=False)
    want = open(os.path.join(golden_dir, "expected", fname + ".k32.w1000.tsv"), "rb").read()
    assert out.read_bytes() == want


@pytest.mark.parametrize("variant", [0, 1, 3, 4])
@pytest.mark.parametrize("canonical", ["sum", "min"])
@pytest.mark.parametrize("fname", FIXTURES)
def test_fixture_vs_oracle(engine, oracle, golden_dir, fname, canonical, variant):
    names, seq, offs = oracle_lib.read_fasta(os.path.join(golden_dir, "inputs", fname))
    engine.set_option("cand_variant", variant)
    for k, w in KW:
        ref = oracle.sketch(seq, offs, k, w, canonical=canonical)
        sk = engine.sketch_buffers(seq, offs, k, w, names=names, canonical=canonical)
        assert_same(sk, ref)
    engine.set_option("cand_variant", DEFAULT_VARIANT)


def _messy(n, seed):
    """random sequence with N runs, lower-case stretches, IUPAC codes, low-complexity blocks, tiny records"""
    rng = np.random.Generator(np.random.PCG64(seed))
    seq = synth.random_bases(n, rng)
    for _ in range(20):
        s = int(rng.integers(0, n - 5000)); ln = int(rng.integers(1, 4000))
        seq[s:s + ln] = ord("N")
    for _ in range(200):
        s = int(rng.integers(0, n - 1)); seq[s] = rng.choice(np.frombuffer(b"NnRYKMSWBDHVU-*", dtype=np.uint8))
    for _ in range(20):
        s = int(rng.integers(0, n - 5000)); ln = int(rng.integers(1, 4000))
        seq[s:s + ln] |= 0x20
    for unit in (b"A", b"AT", b"ACG", b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAC"):
        s = int(rng.integers(0, n - 20000)); ln = int(rng.integers(3000, 12000))
        seq[s:s + ln] = np.resize(np.frombuffer(unit, dtype=np.uint8), ln)
    cuts = np.sort(rng.integers(0, n, size=40))
    offs = np.unique(np.concatenate([[0], cuts, cuts[:5] + 3, cuts[5:8] + 40, [n]])).astype(np.uint64)
    offs = offs[offs <= n]
    offs = np.concatenate([offs[:3], offs[2:3], offs[3:]])   # one empty record
    return seq, offs


@pytest.mark.parametrize("variant", [0, 1, 3, 4])
@pytest.mark.parametrize("canonical", ["sum", "min"])
def test_messy_synthetic(engine, oracle, canonical, variant):
    seq, offs = _messy(1_500_000, 7)
    engine.set_option("cand_variant", variant)
    engine.set_option("prune", variant == 1)     # also exercises the pruned exact path
    for k, w in [(32, 1000), (32, 100), (16, 50), (20, 10), (40, 250), (15, 10), (32, 5000)]:
        ref = oracle.sketch(seq, offs, k, w, canonical=canonical)
        sk = engine.sketch_buffers(seq, offs, k, w, canonical=canonical)
        assert_same(sk, ref)
    engine.set_option("cand_variant", DEFAULT_VARIANT)
    engine.set_option("prune", 0)


@pytest.mark.parametrize("option", ["select_narrow", "fma_offload"])
def test_kernel_variants_behind_options(engine, oracle, option):
    """the 64-bit window selection (inputs beyond 4 G valid k-mers) and the all-ALU threshold test of cand31 are only
    reachable through their options at test sizes: both must give the same sketch as the defaults"""
    seq, offs = _messy(1_200_000, 23)
    engine.set_option(option, 0)
    try:
        for k, w, canonical in [(32, 1000, "sum"), (32, 100, "sum"), (24, 250, "sum"), (40, 500, "min"), (16, 10, "sum")]:
            ref = oracle.sketch(seq, offs, k, w, canonical=canonical)
            assert_same(engine.sketch_buffers(seq, offs, k, w, canonical=canonical), ref)
    finally:
        engine.set_option(option, 1)


@pytest.mark.parametrize("lw", [9, 11, 25, 63])
def test_tile_geometry_independence(engine, oracle, lw):
    """variant 4: the stream length (tile geometry) is a performance knob only -- N runs, record boundaries and the
    sequence end fall at different places inside the tiles for every Lw"""
    seq, offs = _messy(2_100_000, 31)
    engine.set_option("scan_lw", lw)
    try:
        for k, w in [(32, 1000), (40, 100), (24, 250)]:
            ref = oracle.sketch(seq, offs, k, w)
            assert_same(engine.sketch_buffers(seq, offs, k, w), ref)
    finally:
        engine.set_option("scan_lw", 0)


def test_invalid_bytes_every_value(engine, oracle):
    """every byte value once, in every position class of a 32-base unit: the word-parallel validity screen of
    pack2_kernel against the oracle's table (ACGTacgt valid, everything else breaks the k-mers that contain it)"""
    rng = np.random.Generator(np.random.PCG64(5))
    seq = synth.random_bases(256 * 200 + 64, rng)
    for v in range(256):
        seq[200 * v + 37 + (v % 32)] = v
    offs = np.array([0, len(seq)], dtype=np.uint64)
    for k, w in [(32, 20), (24, 10)]:
        ref = oracle.sketch(seq, offs, k, w)
        assert_same(engine.sketch_buffers(seq, offs, k, w), ref)


@pytest.mark.parametrize("tau", [0.5, 3.0, 10.0, 1e9])
def test_threshold_independence(engine, oracle, tau):
    """the candidate threshold is a performance knob only: any tau gives the same sketch"""
    seq, offs = _messy(600_000, 11)
    ref = oracle.sketch(seq, offs, 32, 500)
    engine.set_option("tau", tau)
    try:
        sk = engine.sketch_buffers(seq, offs, 32, 500)
        assert_same(sk, ref)
    finally:
        engine.set_option("tau", 9.0)


def test_config2_scale_down(engine, oracle):
    """BASELINE config 2 generator at 1/10 scale: reference + derived target, k=32 w=1000"""
    rseq, roffs, rnames = synth.make_reference(10_000_000, n_chrom=10)
    tseq, toffs, tnames = synth.derive_target(rseq, roffs)
    for seq, offs in ((rseq, roffs), (tseq, toffs)):
        ref = oracle.sketch(seq, offs, 32, 1000, threads=8)
        sk = engine.sketch_buffers(seq, offs, 32, 1000)
        assert_same(sk, ref)
        c = sk.counts()
        assert c["bases"] == len(seq) and c["minimizers"] == len(ref)


def test_edge_cases(engine, oracle):
    for seq in (b"", b"ACGT", b"A" * 31, b"ACGT" * 8, b"N" * 100, b"ACGT" * 300):
        offs = np.array([0, len(seq)], dtype=np.uint64)
        for k, w in [(32, 1000), (32, 1), (4, 2), (8, 1169)]:
            ref = oracle.sketch(seq, offs, k, w)
            sk = engine.sketch_buffers(seq, offs, k, w)
            assert_same(sk, ref)
    # several empty / short records in a row
    seq = b"ACGTTGCA" * 50
    offs = np.array([0, 0, 10, 10, 45, 400, 400], dtype=np.uint64)
    ref = oracle.sketch(seq, offs, 8, 5)
    assert_same(engine.sketch_buffers(seq, offs, 8, 5), ref)


def test_tsv_with_seq_matches_oracle_cli(engine, golden_dir, tmp_path):
    """`indexlr --seq --long --pos` text (what bin/ntjoin_utils.py:173-185 parses)"""
    import subprocess
    fa = os.path.join(golden_dir, "inputs", "scaf.more_seqs.fa")
    want = subprocess.check_output([oracle_lib.CLI, "--seq", "--long", "--pos", "-k", "32", "-w", "100", fa])
    sk = engine.sketch_file(fa, 32, 100)
    out = tmp_path / "o.tsv"
    sk.write_tsv(out, pos=True, strand=False, seq=True)
    assert out.read_bytes() == want


@pytest.mark.parametrize("variant", [1, 3, 4])
def test_config5_kw_sweep(engine, oracle, variant):
    """BASELINE configs[4]: k in {24,32,40} x w in {250,500,1000,5000}, scaled down, both candidate kernels"""
    rseq, roffs, _ = synth.make_reference(6_000_000, n_chrom=5, dup_frac=0.02, n_frac=0.005)
    engine.set_option("cand_variant", variant)
    try:
        for k in (24, 32, 40):
            for w in (250, 500, 1000, 5000):
                ref = oracle.sketch(rseq, roffs, k, w, threads=8)
                assert_same(engine.sketch_buffers(rseq, roffs, k, w), ref)
    finally:
        engine.set_option("cand_variant", DEFAULT_VARIANT)


def test_multi_assembly_single_call(engine, oracle):
    """several assemblies in one device buffer sketched in ONE call (Engine.sketch_device_multi): every slice must
    equal the sketch of its assembly alone, with and without alignment gaps between the assemblies, and steps 2-3 on
    the slices must equal steps 2-3 on separate sketches"""
    import torch
    a = synth.make_reference(700_000, n_chrom=4, seed=11, dup_frac=0.02, n_frac=0.004)
    b = synth.derive_target(a[0], a[1], min_len=3_000, max_len=90_000)
    c = synth.make_reference(123_457, n_chrom=1, seed=12)
    asms = [(a[0], a[1]), (b[0], b[1]), (c[0], c[1])]
    for k, w in [(32, 100), (12, 7)]:
        refs = [oracle.sketch(s, o, k, w) for s, o in asms]
        for aligned in (False, True):
            starts, at = [], 0
            for s, _ in asms:
                starts.append(at)
                at += (len(s) + 15) // 16 * 16 if aligned else len(s)
            buf = np.full(at + 16, ord("A"), dtype=np.uint8)          # gap bytes are valid bases: the gap record must not leak
            for (s, _), st in zip(asms, starts):
                buf[st:st + len(s)] = s
            dbuf = torch.from_numpy(buf).cuda()
            parent, slices = engine.sketch_device_multi(dbuf.data_ptr(), [o for _, o in asms], k, w, starts=starts)
            for sl, ref in zip(slices, refs):
                assert_same(sl, ref)
            res = engine.filter_and_edges(slices, [2.0, 2.0, 1.0])
            want = oracle.filter_and_edges([r["out_hash"] for r in refs], [r["contig"] for r in refs], [2.0, 2.0, 1.0])
            np.testing.assert_array_equal(res.vertices, want["vertices"])
            np.testing.assert_array_equal(res.edge_u, want["edges"]["u"])
            np.testing.assert_array_equal(res.edge_v, want["edges"]["v"])
            np.testing.assert_array_equal(res.support, want["edges"]["support_mask"])
            res.close()
            parent.close()


@pytest.mark.parametrize("scale", [0.002, 0.05])
def test_size_bounds_overflow_repeats_exactly(engine, oracle, scale):
    """the sketch sizes its candidate / gap / minimizer arrays from bounds and reads the exact counts once, at the end;
    when a bound is too small (forced here by shrinking them) the call repeats itself with exact sizes -- same result,
    from device and from host buffers, and the same as with the round trips of the exact path"""
    import torch
    seq, offs = _messy(900_000, 41)
    ref = oracle.sketch(seq, offs, 32, 100)
    engine.set_option("bound_scale", scale)
    try:
        assert_same(engine.sketch_buffers(seq, offs, 32, 100), ref)
        dseq = torch.from_numpy(np.ascontiguousarray(seq)).cuda()
        assert_same(engine.sketch_device(dseq.data_ptr(), offs, 32, 100), ref)
    finally:
        engine.set_option("bound_scale", 1.0)
    engine.set_option("async_sizes", 0)
    try:
        assert_same(engine.sketch_buffers(seq, offs, 32, 100), ref)
    finally:
        engine.set_option("async_sizes", 1)


def test_concurrent_sketches_two_streams(engine, oracle):
    """Engine.sketch_device_many: several assemblies enqueued on two streams; same tuples as one call each, and the
    filter that follows on the engine stream sees both results"""
    import torch
    a = synth.make_reference(2_000_000, n_chrom=4, seed=21, dup_frac=0.02, n_frac=0.004)
    b = synth.derive_target(a[0], a[1], min_len=3_000, max_len=200_000)
    c = synth.make_reference(300_001, n_chrom=2, seed=22)
    asms = [(a[0], a[1]), (b[0], b[1]), (c[0], c[1])]
    refs = [oracle.sketch(s, o, 32, 250) for s, o in asms]
    dev = [torch.from_numpy(np.ascontiguousarray(s)).cuda() for s, _ in asms]
    for rep in range(3):
        sks = engine.sketch_device_many([d.data_ptr() for d in dev], [o for _, o in asms], 32, 250)
        res = engine.filter_and_edges(sks, [2.0, 1.0, 1.0])
        for sk, ref in zip(sks, refs):
            assert_same(sk, ref)
        want = oracle.filter_and_edges([r["out_hash"] for r in refs], [r["contig"] for r in refs], [2.0, 1.0, 1.0])
        np.testing.assert_array_equal(res.vertices, want["vertices"])
        np.testing.assert_array_equal(res.edge_u, want["edges"]["u"])
        res.close()
        for sk in sks:
            sk.close()
    # a too-small size bound on one of them: that assembly repeats itself with exact sizes
    engine.set_option("bound_scale", 0.01)
    try:
        sks = engine.sketch_device_many([d.data_ptr() for d in dev], [o for _, o in asms], 32, 250)
        for sk, ref in zip(sks, refs):
            assert_same(sk, ref)
    finally:
        engine.set_option("bound_scale", 1.0)


def test_host_copy_started_early(engine, oracle):
    """Engine.sketch_many(prefetch_host=True) / Sketch.prefetch_host(): the device->host copy of the tuples is started on
    the engine's copy stream as soon as a sketch is done (beside the next assembly's host->device copy and steps 2-3);
    views and copies taken afterwards hold the same tuples, a sketch closed with the copy in flight is released cleanly"""
    a = synth.make_reference(1_500_000, n_chrom=3, seed=31, dup_frac=0.02, n_frac=0.004)
    b = synth.derive_target(a[0], a[1], min_len=3_000, max_len=200_000)
    refs = [oracle.sketch(s, o, 32, 100) for s, o in ((a[0], a[1]), (b[0], b[1]))]
    for rep in range(3):                                   # the pinned blocks are pooled: reuse them
        sks = engine.sketch_many([(a[0], a[1]), (b[0], b[1])], 32, 100, prefetch_host=True)
        res = engine.filter_and_edges(sks, [2.0, 1.0])     # runs on the engine stream while the copies are in flight
        for sk, ref in zip(sks, refs):
            views = sk.fetch(copy=False)
            np.testing.assert_array_equal(views[0], ref["out_hash"])
            np.testing.assert_array_equal(views[2].astype(np.uint64), ref["pos"])
            assert_same(sk, ref)
        want = oracle.filter_and_edges([r["out_hash"] for r in refs], [r["contig"] for r in refs], [2.0, 1.0])
        np.testing.assert_array_equal(res.edge_u, want["edges"]["u"])
        res.close()
        for sk in sks:
            sk.close()
    sk = engine.sketch_buffers(a[0], a[1], 32, 100).prefetch_host()
    sk.prefetch_host()                                     # second call: nothing to do
    sk.close()                                             # closed with the copy possibly still in flight
    sk = engine.sketch_buffers(a[0], a[1], 32, 100)
    assert_same(sk, refs[0])                               # plain lazy path still works
    sk.prefetch_host()                                     # after the host copy exists: no-op
    sk.close()
