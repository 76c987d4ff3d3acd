"""btllib-shaped Python objects backed by the engine.

ntJoin uses two btllib classes from Python after the hot path:
  * btllib.Indexlr(path, k, w, IndexlrFlag.LONG_MODE, threads)  -- overlap re-sketch at k=15, w=10 on
    <prefix>.segments.fa; consumers read record.id and record.minimizers[i].out_hash / .pos
    (bin/ntjoin_assemble.py:478-481, 506-507)
  * btllib.SeqReader(path, SeqReaderFlag.LONG_MODE, threads)     -- record.id / record.seq
    (bin/ntjoin_assemble.py:313-316)
These stand-ins expose the same attribute names so that `import ntjoin_b200.btllib_compat as btllib` works for
those call sites; the minimizers come from the same CUDA kernels as step 1.
"""
import os
from collections import namedtuple

Minimizer = namedtuple("Minimizer", ["min_hash", "out_hash", "pos", "forward", "seq"])
IndexlrRecord = namedtuple("IndexlrRecord", ["num", "id", "barcode", "readlen", "minimizers"])
SeqRecord = namedtuple("SeqRecord", ["num", "id", "comment", "seq", "qual"])


class IndexlrFlag:
    NO_ID, BX, SEQ, FILTER_IN, FILTER_OUT, SHORT_MODE, LONG_MODE = 1, 2, 4, 8, 16, 32, 64


class SeqReaderFlag:
    FOLD_CASE, NO_FOLD_CASE, NO_TRIM_MASKED, TRIM_MASKED, SHORT_MODE, LONG_MODE = 0, 1, 0, 2, 4, 8


_ENGINE = None


def _engine():
    global _ENGINE
    if _ENGINE is None:
        from .engine import Engine
        _ENGINE = Engine(int(os.environ.get("MXE_DEVICE", "0")))
    return _ENGINE


class Indexlr:
    """Iterates records of a FASTA file with their ordered minimizers (context manager like btllib's)."""

    def __init__(self, seqfile, k, w, flags=IndexlrFlag.LONG_MODE, threads=5, verbose=False, canonical="sum"):
        if not os.path.exists(seqfile):
            raise FileNotFoundError(seqfile)
        self._sk = _engine().sketch_file(seqfile, k, w, canonical=canonical)
        self._k, self._flags = k, flags

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        if self._sk is not None:
            self._sk.close()
            self._sk = None

    def __iter__(self):
        sk = self._sk
        oh, mh, ps, fw = sk.out_hash, sk.min_hash, sk.pos, sk.forward
        cg = sk.contig
        import numpy as np
        bounds = np.searchsorted(cg, np.arange(len(sk.names) + 1))
        for c, name in enumerate(sk.names):
            s, e = int(bounds[c]), int(bounds[c + 1])
            mxs = [Minimizer(int(mh[i]), int(oh[i]), int(ps[i]), bool(fw[i]), "") for i in range(s, e)]
            yield IndexlrRecord(c, name, "", 0, mxs)


class SeqReader:
    """FASTA/FASTQ records with upper-cased sequence (btllib folds case by default)."""

    def __init__(self, seqfile, flags=SeqReaderFlag.LONG_MODE, threads=5):
        if not os.path.exists(seqfile):
            raise FileNotFoundError(seqfile)
        self._path = seqfile

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass

    def __iter__(self):
        # the engine's own reader (host only: mxe_fasta_read), the one mxe_sketch_file sketches from
        import ctypes as C
        from ._lib import check, load_library
        lib = load_library()
        h = C.c_void_p()
        check(lib, lib.mxe_fasta_read(str(self._path).encode(), C.byref(h)))
        try:
            n, offs, seq, nm = C.c_uint32(), C.c_void_p(), C.c_void_p(), C.c_char_p()
            check(lib, lib.mxe_fasta_view(h, C.byref(n), C.byref(offs), C.byref(seq)))
            offsets = C.cast(offs, C.POINTER(C.c_uint64))
            for i in range(n.value):
                check(lib, lib.mxe_fasta_name(h, i, C.byref(nm)))
                a, b = offsets[i], offsets[i + 1]
                yield SeqRecord(i, nm.value.decode("utf-8", "replace"), "", C.string_at(seq.value + a, b - a).decode("ascii", "replace"), "")
        finally:
            lib.mxe_fasta_free(h)


