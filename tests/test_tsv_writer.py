"""CPU: the `indexlr` TSV text (mxe_write_tsv over a host-only sketch object, mxe_sketch_from_arrays) byte-equal to the
oracle CLI for every flag combination ntJoin and the goldens use, whatever the number of formatting threads."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
import ntjoin_b200
from ntjoin_b200._lib import check
from ntjoin_b200 import synth


def _write(m, names, seq, offs, k, path, pos, strand, with_seq):
    lib = ntjoin_b200.load_library()
    oh = np.ascontiguousarray(m["out_hash"], dtype=np.uint64)
    mh = np.ascontiguousarray(m["min_hash"], dtype=np.uint64)
    ps = np.ascontiguousarray(m["pos"], dtype=np.uint32)
    cg = np.ascontiguousarray(m["contig"], dtype=np.uint32)
    fw = np.ascontiguousarray(m["forward"], dtype=np.uint8)
    of = np.ascontiguousarray(offs, dtype=np.uint64)
    sq = np.ascontiguousarray(np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else seq, dtype=np.uint8)
    nm = (C.c_char_p * len(names))(*[n.encode() for n in names])
    h = C.c_void_p()
    check(lib, lib.mxe_sketch_from_arrays(oh.ctypes.data, mh.ctypes.data, ps.ctypes.data, cg.ctypes.data, fw.ctypes.data, len(oh), nm,
                                          of.ctypes.data_as(C.POINTER(C.c_uint64)), len(names), k, sq.ctypes.data, C.byref(h)))
    try:
        check(lib, lib.mxe_write_tsv(h, str(path).encode(), int(pos), int(strand), int(with_seq)))
    finally:
        lib.mxe_sketch_free(h)


@pytest.mark.parametrize("threads", ["1", "4"])
def test_fixture_tsv_bytes(oracle, golden_dir, tmp_path, monkeypatch, threads):
    monkeypatch.setenv("MXE_HOST_THREADS", threads)
    for f, k, w in [("ref.fa", 32, 500), ("scaf.multiple.fa", 32, 250), ("scaf.more_seqs.fa", 15, 10), ("scaf.f-f.termN.unassigned.fa", 24, 100)]:
        fa = os.path.join(golden_dir, "inputs", f)
        names, seq, offs = oracle_lib.read_fasta(fa)
        m = oracle.sketch(seq, offs, k, w)
        for flags, (pos, strand, wseq) in {("--pos", "--seq"): (1, 0, 1), ("--pos",): (1, 0, 0), ("--pos", "--strand", "--seq"): (1, 1, 1), (): (0, 0, 0)}.items():
            want = subprocess.check_output([oracle_lib.CLI, *flags, "--long", "-k", str(k), "-w", str(w), fa])
            out = tmp_path / "o.tsv"
            _write(m, names, seq, offs, k, out, pos, strand, wseq)
            assert out.read_bytes() == want, (f, flags)


@pytest.mark.parametrize("threads", ["1", "2", "3", "7", "16"])
def test_large_tsv_is_thread_count_independent(oracle, tmp_path, monkeypatch, threads):
    """> 65536 minimizers (the multi-threaded path), many records incl. records without minimizers and lower-case sequence"""
    rseq, roffs, rnames = synth.make_reference(3_000_000, n_chrom=5, dup_frac=0.02, n_frac=0.01, seed=9)
    seq, offs, names = synth.derive_target(rseq, roffs, min_len=40, max_len=60_000, seed=10)
    seq = seq.copy()
    seq[1000:400000] |= 0x20
    fa = tmp_path / "t.fa"
    synth.write_fasta(fa, seq, offs, names)
    m = oracle.sketch(seq, offs, 32, 20)
    assert len(m) > 1 << 17 and len(set(m["contig"].tolist())) < len(names)
    want = subprocess.check_output([oracle_lib.CLI, "--seq", "--long", "--pos", "-k", "32", "-w", "20", str(fa)])
    monkeypatch.setenv("MXE_HOST_THREADS", threads)
    out = tmp_path / "o.tsv"
    _write(m, names, seq, offs, 32, out, 1, 0, 1)
    assert out.read_bytes() == want
