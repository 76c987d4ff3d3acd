"""CPU: the engine's FASTA/FASTQ reader on its own (mxe_fasta_read, host only -- the reader mxe_sketch_file sketches from
and btllib_compat.SeqReader serves) against the test-side Python reader."""
import gzip
import os

import pytest

import oracle_lib
from ntjoin_b200 import btllib_compat as btllib


def _records(path):
    with btllib.SeqReader(str(path), btllib.SeqReaderFlag.LONG_MODE, 2) as rd:
        return [(r.id, r.seq) for r in rd]


def test_fixtures_match_python_reader(golden_dir):
    for f in sorted(os.listdir(os.path.join(golden_dir, "inputs"))):
        path = os.path.join(golden_dir, "inputs", f)
        names, seq, offs = oracle_lib.read_fasta(path)
        want = [(n, bytes(seq[int(offs[i]):int(offs[i + 1])]).decode()) for i, n in enumerate(names)]
        assert _records(path) == want, f


def test_messy_fasta(tmp_path):
    p = tmp_path / "m.fa"
    p.write_bytes(b"ignored line before the first header\n>r1 some description\tmore\nACGT\nacgtnn\n\n>r2\r\nGG\r\nTT\r\n>empty\n>r3|x\nNNNN\n>last\nAC")
    assert _records(p) == [("r1", "ACGTACGTNN"), ("r2", "GGTT"), ("empty", ""), ("r3|x", "NNNN"), ("last", "AC")]


def test_fastq_and_gz(tmp_path):
    p = tmp_path / "r.fq"
    p.write_text("@a desc\nACGTn\n+\nIIIII\n@b\nGG\n+b\nII\n")
    assert _records(p) == [("a", "ACGTN"), ("b", "GG")]
    z = tmp_path / "z.fa.gz"
    with gzip.open(z, "wt") as fh:
        fh.write(">g1 c\nacgt\nAC\n>g2\nT\n")
    assert _records(z) == [("g1", "ACGTAC"), ("g2", "T")]


def test_missing_file_raises(tmp_path):
    with pytest.raises(FileNotFoundError):
        btllib.SeqReader(str(tmp_path / "nope.fa"), btllib.SeqReaderFlag.LONG_MODE, 1)


def _big_messy_fasta(path, seed=3, n_records=40):
    """about 28 MB, four 8 MB chunks: multi-line records of mixed line widths, CRLF stretches, lower case, blank lines,
    headers with descriptions, empty records, junk before the first header, no newline at the end"""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    alphabet = np.frombuffer(b"ACGTacgtNnRY", dtype=np.uint8)
    want, out = [], [b"this line and the next come before any header\nACGT\n"]
    for r in range(n_records):
        name = f"rec{r}|x"
        n = int(rng.choice([0, 1, 59, 60, 61, 5_000, 300_000, 3_000_000]))
        seq = alphabet[rng.integers(0, len(alphabet), n)].tobytes()
        want.append((name, seq.decode().upper()))
        out.append(f">{name} description {r}\tmore".encode() + (b"\r\n" if r % 3 == 0 else b"\n"))
        width = int(rng.choice([60, 61, 80, 1000, 10_000_000]))
        eol = b"\r\n" if r % 3 == 0 else b"\n"
        for i in range(0, n, width):
            out.append(seq[i:i + width] + eol)
            if r % 7 == 0 and i == 0:
                out.append(eol)                      # a blank line inside a record
    data = b"".join(out)
    data = data[:-1] if data.endswith(b"\n") and not data.endswith(b"\r\n") else data
    with open(path, "wb") as fh:
        fh.write(data)
    return want


@pytest.mark.parametrize("threads", ["1", "2", "3", "7", "16"])
def test_parallel_reader_is_thread_count_independent(tmp_path, monkeypatch, threads):
    p = tmp_path / "big.fa"
    want = _big_messy_fasta(p)
    assert os.path.getsize(p) > 24 << 20          # at least four 8 MB chunks
    monkeypatch.setenv("MXE_HOST_THREADS", threads)
    got = _records(p)
    assert [g[0] for g in got] == [w[0] for w in want]
    assert all(g[1] == w[1] for g, w in zip(got, want))


def test_reader_edge_files(tmp_path):
    cases = {
        "empty.fa": (b"", []),
        "only_junk.fa": (b"no header here\nACGT\n", []),
        "header_only.fa": (b">x", [("x", "")]),
        "one_line_no_nl.fa": (b">x y\nacgt", [("x", "ACGT")]),
        "crlf_end.fa": (b">a\r\nAC\r\n>b\r\n\r\nGT\r\n", [("a", "AC"), ("b", "GT")]),
        "gt_inside.fa": (b">a\nAC>GT\n", [("a", "AC>GT")]),
    }
    for name, (data, want) in cases.items():
        p = tmp_path / name
        p.write_bytes(data)
        assert _records(p) == want, name


def _read_c(path):
    import ctypes as C
    from ntjoin_b200._lib import load_library
    lib = load_library()
    h = C.c_void_p()
    rc = lib.mxe_fasta_read(str(path).encode(), C.byref(h))
    if rc != 0:
        return rc, lib.mxe_last_error()
    n, offs, text, nm = C.c_uint32(), C.c_void_p(), C.c_void_p(), C.c_char_p()
    assert lib.mxe_fasta_view(h, C.byref(n), C.byref(offs), C.byref(text)) == 0
    o = C.cast(offs, C.POINTER(C.c_uint64))
    out = []
    for i in range(n.value):
        assert lib.mxe_fasta_name(h, i, C.byref(nm)) == 0
        out.append((nm.value.decode(), C.string_at(text.value + o[i], o[i + 1] - o[i]).decode()))
    lib.mxe_fasta_free(h)
    return 0, out


def test_compressed_input_is_read_through_the_decompressor(tmp_path):
    """btllib reads gzip / bzip2 / xz / zstd input through a pipe from the external tool; so does mxe_fasta_read (by the
    magic bytes, whatever the file is called).  A damaged file is an error (MXE_ERR_IO), never a sketch of compressed bytes."""
    import bz2
    import gzip
    import lzma
    import shutil
    text = b">a d\nACGTacgtNN\nGG\n>b\n" + b"ACGT" * 50000 + b"\n"
    want = [("a", "ACGTACGTNNGG"), ("b", "ACGT" * 50000)]
    for tool, name, opener in (("gzip", "x.fa.gz", gzip.open), ("bzip2", "y.fa.bz2", bz2.open), ("xz", "z.data", lzma.open)):
        p = tmp_path / name
        with opener(p, "wb") as fh:
            fh.write(text)
        if shutil.which(tool) is None:
            continue
        assert _read_c(p) == (0, want), tool
    fq = tmp_path / "r.fq.gz"
    with gzip.open(fq, "wb") as fh:
        fh.write(b"@r1 d\nACgt\n+\nIIII\n")
    assert _read_c(fq) == (0, [("r1", "ACGT")])
    quoted = tmp_path / "it's here.fa.gz"                      # the path goes through a shell command line
    with gzip.open(quoted, "wb") as fh:
        fh.write(b">q\nAC\n")
    assert _read_c(quoted) == (0, [("q", "AC")])
    bad = tmp_path / "bad.fa.gz"
    bad.write_bytes((tmp_path / "x.fa.gz").read_bytes()[:40])    # truncated stream
    rc, msg = _read_c(bad)
    assert rc == -2 and b"failed" in msg


def test_fastq_trailing_blank_line(tmp_path):
    import ctypes as C
    from ntjoin_b200._lib import load_library
    lib = load_library()
    p = tmp_path / "x.fq"
    p.write_bytes(b"@r1 d\nACGT\n+\nIIII\n@r2\nGGCC\n+\nIIII\n\n")
    h = C.c_void_p()
    assert lib.mxe_fasta_read(str(p).encode(), C.byref(h)) == 0
    n = C.c_uint32()
    offs, text = C.c_void_p(), C.c_void_p()
    assert lib.mxe_fasta_view(h, C.byref(n), C.byref(offs), C.byref(text)) == 0
    assert n.value == 2
    lib.mxe_fasta_free(h)


def test_compressed_big_messy_file_equals_plain(tmp_path):
    """the ~28 MB messy file through `gzip -dc` gives the same records as the mapped plain file; a second file whose text
    is larger than the first block of the pipe buffer (64 MB) exercises the buffer growth"""
    import shutil
    if shutil.which("gzip") is None:
        pytest.skip("gzip not installed")
    p = tmp_path / "big.fa"
    want = _big_messy_fasta(p)
    z = tmp_path / "big.fa.gz"
    with open(p, "rb") as src, gzip.open(z, "wb", compresslevel=1) as dst:
        shutil.copyfileobj(src, dst)
    got = _records(z)
    assert [g[0] for g in got] == [w[0] for w in want] and all(g[1] == w[1] for g, w in zip(got, want))
    long_rec = b"ACGTTGCAAC" * 9_000_000                     # 90 MB of sequence in lines of 100
    z2 = tmp_path / "long.fa.gz"
    with gzip.open(z2, "wb", compresslevel=1) as dst:
        dst.write(b">only one\n")
        for i in range(0, len(long_rec), 9_000_000):
            block = long_rec[i:i + 9_000_000]
            dst.write(b"\n".join(block[j:j + 100] for j in range(0, len(block), 100)) + b"\n")
    rc, recs = _read_c(z2)
    assert rc == 0 and len(recs) == 1 and recs[0][0] == "only" and len(recs[0][1]) == len(long_rec)
    assert recs[0][1][:20] == long_rec[:20].decode() and recs[0][1][-20:] == long_rec[-20:].decode()
