"""ctypes face of the CPU oracle (oracle/mxo.c).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_build", "libmxo.so")
CLI = os.path.join(ORACLE_DIR, "_build", "mxo_indexlr")


def build():
    if not (os.path.exists(LIB) and os.path.exists(CLI)) or \
            os.path.getmtime(LIB) < os.path.getmtime(os.path.join(ORACLE_DIR, "mxo.c")):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])
    return LIB


class MxoMx(C.Structure):
    _fields_ = [("out_hash", C.c_uint64), ("min_hash", C.c_uint64), ("pos", C.c_uint64),
                ("contig", C.c_uint32), ("forward", C.c_uint32)]


class MxoEdge(C.Structure):
    _fields_ = [("u", C.c_uint64), ("v", C.c_uint64), ("support_mask", C.c_uint32), ("weight", C.c_double)]


MX_DTYPE = np.dtype([("out_hash", "<u8"), ("min_hash", "<u8"), ("pos", "<u8"), ("contig", "<u4"), ("forward", "<u4")])
EDGE_DTYPE = np.dtype([("u", "<u8"), ("v", "<u8"), ("support_mask", "<u4"), ("_pad", "<u4"), ("weight", "<f8")])


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build())
        L = self.lib
        L.mxo_sketch_buffers.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_uint32, C.c_uint, C.c_uint, C.c_int, C.c_int,
                                         C.c_int, C.POINTER(C.POINTER(MxoMx)), C.POINTER(C.c_size_t)]
        L.mxo_free.argtypes = [C.c_void_p]
        L.mxo_free.restype = None
        L.mxo_kmer_hashes.argtypes = [C.c_char_p, C.c_uint, C.c_int] + [C.POINTER(C.c_uint64)] * 4
        L.mxo_kmer_hashes.restype = None
        L.mxo_filter_and_edges.argtypes = [C.c_int, C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.POINTER(C.c_uint32)),
                                           C.POINTER(C.c_size_t), C.POINTER(C.c_double),
                                           C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.POINTER(C.c_uint8)),
                                           C.POINTER(C.POINTER(MxoEdge)), C.POINTER(C.c_size_t),
                                           C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_size_t)]

    def sketch(self, seq, offsets, k, w, canonical="sum", tie="right", threads=1):
        """seq: bytes / numpy uint8 (concatenated records); returns structured array MX_DTYPE."""
        a = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
        offs = np.ascontiguousarray(offsets, dtype=np.uint64)
        out = C.POINTER(MxoMx)()
        n = C.c_size_t()
        rc = self.lib.mxo_sketch_buffers(a.ctypes.data, offs.ctypes.data_as(C.POINTER(C.c_uint64)), len(offs) - 1, k, w,
                                         {"sum": 0, "min": 1}[canonical], {"right": 0, "left": 1}[tie], threads,
                                         C.byref(out), C.byref(n))
        if rc:
            raise RuntimeError(f"oracle sketch failed: {rc}")
        res = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint8)), shape=(n.value * C.sizeof(MxoMx),)).view(MX_DTYPE).copy() \
            if n.value else np.empty(0, dtype=MX_DTYPE)
        self.lib.mxo_free(out)
        return res

    def kmer_hashes(self, kmer, canonical="sum"):
        v = [C.c_uint64() for _ in range(4)]
        self.lib.mxo_kmer_hashes(kmer if isinstance(kmer, bytes) else kmer.encode(), len(kmer), {"sum": 0, "min": 1}[canonical],
                                 *[C.byref(x) for x in v])
        return tuple(x.value for x in v)   # fwd, rev, hash0, hash1

    def filter_and_edges(self, hashes, contigs, weights):
        n_asm = len(hashes)
        hs = [np.ascontiguousarray(h, dtype=np.uint64) for h in hashes]
        cs = [np.ascontiguousarray(c, dtype=np.uint32) for c in contigs]
        uniq = [np.zeros(len(h), dtype=np.uint8) for h in hs]
        keep = [np.zeros(len(h), dtype=np.uint8) for h in hs]
        HP = (C.POINTER(C.c_uint64) * n_asm)(*[h.ctypes.data_as(C.POINTER(C.c_uint64)) for h in hs])
        CP = (C.POINTER(C.c_uint32) * n_asm)(*[c.ctypes.data_as(C.POINTER(C.c_uint32)) for c in cs])
        UP = (C.POINTER(C.c_uint8) * n_asm)(*[u.ctypes.data_as(C.POINTER(C.c_uint8)) for u in uniq])
        KP = (C.POINTER(C.c_uint8) * n_asm)(*[k.ctypes.data_as(C.POINTER(C.c_uint8)) for k in keep])
        NN = (C.c_size_t * n_asm)(*[len(h) for h in hs])
        WW = (C.c_double * n_asm)(*[float(x) for x in weights])
        E = C.POINTER(MxoEdge)()
        ne = C.c_size_t()
        V = C.POINTER(C.c_uint64)()
        nv = C.c_size_t()
        rc = self.lib.mxo_filter_and_edges(n_asm, HP, CP, NN, WW, UP, KP, C.byref(E), C.byref(ne), C.byref(V), C.byref(nv))
        if rc:
            raise RuntimeError(f"oracle filter failed: {rc}")
        edges = np.ctypeslib.as_array(C.cast(E, C.POINTER(C.c_uint8)), shape=(ne.value * C.sizeof(MxoEdge),)).view(EDGE_DTYPE).copy() \
            if ne.value else np.empty(0, dtype=EDGE_DTYPE)
        edges["_pad"] = 0     # struct padding is uninitialised on the C side
        verts = np.ctypeslib.as_array(V, shape=(nv.value,)).copy() if nv.value else np.empty(0, dtype=np.uint64)
        self.lib.mxo_free(E)
        self.lib.mxo_free(V)
        return {"uniq": [u.astype(bool) for u in uniq], "keep": [k.astype(bool) for k in keep], "edges": edges, "vertices": verts}


def read_fasta(path):
    """Small pure-Python FASTA reader for tests: (names, seq bytes upper-cased, offsets)."""
    import gzip
    op = gzip.open if str(path).endswith(".gz") else open
    names, parts, offsets, at = [], [], [0], 0
    cur = []
    with op(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if names:
                    s = b"".join(cur).upper()
                    parts.append(s)
                    at += len(s)
                    offsets.append(at)
                names.append(line[1:].split()[0].decode() if line[1:].split() else "")
                cur = []
            else:
                cur.append(line.strip())
    if names:
        s = b"".join(cur).upper()
        parts.append(s)
        at += len(s)
        offsets.append(at)
    return names, b"".join(parts), np.array(offsets, dtype=np.uint64)
