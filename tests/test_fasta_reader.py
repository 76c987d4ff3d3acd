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
