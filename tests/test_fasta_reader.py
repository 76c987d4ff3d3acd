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


def test_compressed_input_is_refused(tmp_path):
    """btllib reads .gz transparently; this reader takes plain text and must say so (MXE_ERR_IO) rather than sketch
    compressed bytes as bases and exit 0"""
    import ctypes as C
    import gzip
    from ntjoin_b200._lib import load_library
    lib = load_library()
    p = tmp_path / "x.fa.gz"
    with gzip.open(p, "wb") as fh:
        fh.write(b">a\nACGTACGTACGT\n")
    h = C.c_void_p()
    rc = lib.mxe_fasta_read(str(p).encode(), C.byref(h))
    assert rc == -2 and b"compressed" in lib.mxe_last_error()


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
