"""btllib for the reference harness: SeqReader and Indexlr (bin/ntjoin_assemble.py:313-316, :478-481, :506-507) served by
the CPU oracle (oracle/mxo.c through tests/oracle_lib.py), so that the overlap re-sketch of the reference's own tests
(k=15, w=10 on N-masked segments) is computed by the same restatement as the `indexlr` stand-in."""
import os
import sys
from collections import namedtuple

_TESTS = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
if _TESTS not in sys.path:
    sys.path.insert(0, _TESTS)
import oracle_lib  # noqa: E402

Minimizer = namedtuple("Minimizer", ["min_hash", "out_hash", "pos", "forward", "seq"])
IndexlrRecord = namedtuple("IndexlrRecord", ["num", "id", "barcode", "readlen", "minimizers"])
SeqRecord = namedtuple("SeqRecord", ["num", "id", "comment", "seq", "qual"])


class IndexlrFlag:
    NO_ID, BX, SEQ, FILTER_IN, FILTER_OUT, SHORT_MODE, LONG_MODE = 1, 2, 4, 8, 16, 32, 64


class SeqReaderFlag:
    FOLD_CASE, NO_FOLD_CASE, NO_TRIM_MASKED, TRIM_MASKED, SHORT_MODE, LONG_MODE = 0, 1, 0, 2, 4, 8


class _Ctx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass


class SeqReader(_Ctx):
    def __init__(self, seqfile, flags=SeqReaderFlag.LONG_MODE, threads=5):
        if not os.path.exists(seqfile):
            raise FileNotFoundError(seqfile)
        self._path = seqfile

    def __iter__(self):
        names, seq, offs = oracle_lib.read_fasta(self._path)       # upper-cased, headers cut at the first blank
        for i, name in enumerate(names):
            yield SeqRecord(i, name, "", bytes(seq[int(offs[i]):int(offs[i + 1])]).decode(), "")


class Indexlr(_Ctx):
    _oracle = None

    def __init__(self, seqfile, k, w, flags=IndexlrFlag.LONG_MODE, threads=5, verbose=False):
        if not os.path.exists(seqfile):
            raise FileNotFoundError(seqfile)
        if Indexlr._oracle is None:
            Indexlr._oracle = oracle_lib.Oracle()
        self._path, self._k, self._w = seqfile, k, w

    def __iter__(self):
        names, seq, offs = oracle_lib.read_fasta(self._path)
        m = Indexlr._oracle.sketch(seq, offs, self._k, self._w, canonical=os.environ.get("MXO_CANONICAL") or "sum")
        at = 0
        for i, name in enumerate(names):                           # every record, with or without minimizers
            mxs = []
            while at < len(m) and int(m["contig"][at]) == i:
                mxs.append(Minimizer(int(m["min_hash"][at]), int(m["out_hash"][at]), int(m["pos"][at]), bool(m["forward"][at]), ""))
                at += 1
            yield IndexlrRecord(i, name, "", int(offs[i + 1] - offs[i]), mxs)


if os.environ.get("NTJOIN_B200", "0") not in ("", "0") and os.environ.get("MXE_REPO_ROOT"):
    # drop-in run: the PRODUCT's SeqReader (host-only C reader, mxe_fasta_read) serves bin/ntjoin_assemble.py:313-316
    if os.environ["MXE_REPO_ROOT"] not in sys.path:
        sys.path.insert(0, os.environ["MXE_REPO_ROOT"])
    from ntjoin_b200.btllib_compat import SeqReader, SeqReaderFlag  # noqa: E402,F401,F811
