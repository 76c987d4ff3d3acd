"""Python face of the engine: thin objects over the C ABI (include/mxe.h).

Reference interfaces mirrored here:
  Engine.sketch_file      <-> `indexlr --seq --long --pos -k K -w W -t T FASTA` (ntJoin:204-205)
                              and btllib.Indexlr(path, k, w, ...) (bin/ntjoin_assemble.py:478-481)
  Engine.filter_and_edges <-> read_minimizers uniqueness + filter_minimizers + build_graph edge stage
                              (bin/ntjoin_utils.py:167-193, :152-165, :94-115)
"""
import ctypes as C
import weakref

import numpy as np

from ._lib import check, load_library

CANON = {"sum": 0, "min": 1}


def _np_view(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.empty(0, dtype=dtype)
    ct = np.ctypeslib.as_ctypes_type(dtype)
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,))
    return arr


class Sketch:
    """Ordered minimizers of one assembly: SoA sorted by (contig, pos), contigs in input order."""

    def __init__(self, engine, handle, names, keepalive=None):
        self._e, self._h, self.names = engine, handle, list(names)
        self._keep = keepalive
        self._views = None
        engine._live.add(self)

    def _view(self):
        if self._views is None:
            lib = self._e._lib
            n = C.c_uint64()
            p = [C.c_void_p() for _ in range(5)]
            check(lib, lib.mxe_sketch_view(self._h, C.byref(n), *[C.byref(x) for x in p]))
            n = n.value
            dt = [np.uint64, np.uint64, np.uint32, np.uint32, np.uint8]
            self._views = tuple(_np_view(x.value, n, d).copy() for x, d in zip(p, dt))
        return self._views

    def fetch(self, copy=True):
        """(out_hash, min_hash, pos, contig, forward) in host memory: one pinned device->host copy on first use.
        copy=False returns views into the engine's pinned block (valid until close()) instead of numpy copies."""
        if copy:
            return self._view()
        lib = self._e._lib
        n = C.c_uint64()
        p = [C.c_void_p() for _ in range(5)]
        check(lib, lib.mxe_sketch_view(self._h, C.byref(n), *[C.byref(x) for x in p]))
        dt = [np.uint64, np.uint64, np.uint32, np.uint32, np.uint8]
        return tuple(_np_view(x.value, n.value, d) for x, d in zip(p, dt))

    def prefetch_host(self):
        """Start the device->host copy of the tuples now and return at once (mxe_sketch_prefetch_host): it runs beside the
        next assembly's host->device copy and sketch; a later fetch() / view only waits for it."""
        check(self._e._lib, self._e._lib.mxe_sketch_prefetch_host(self._h))
        return self

    out_hash = property(lambda s: s._view()[0])
    min_hash = property(lambda s: s._view()[1])
    pos = property(lambda s: s._view()[2])
    contig = property(lambda s: s._view()[3])
    forward = property(lambda s: s._view()[4])

    @property
    def n(self):
        lib = self._e._lib
        n = C.c_uint64()
        check(lib, lib.mxe_sketch_device_view(self._h, C.byref(n), None, None, None))
        return n.value

    def device_pointers(self):
        """(n, d_out_hash, d_pos, d_contig) raw device addresses (uint64 / uint32 / uint32 arrays)."""
        lib = self._e._lib
        n = C.c_uint64()
        p = [C.c_void_p() for _ in range(3)]
        check(lib, lib.mxe_sketch_device_view(self._h, C.byref(n), *[C.byref(x) for x in p]))
        return (n.value,) + tuple(x.value or 0 for x in p)

    def counts_raw(self):
        lib = self._e._lib
        a = [C.c_uint64() for _ in range(4)]
        c = C.c_uint32()
        check(lib, lib.mxe_sketch_counts(self._h, *[C.byref(x) for x in a], C.byref(c)))
        return a[0].value, a[1].value, a[2].value, a[3].value, c.value

    def counts(self):
        b, v, cd, g, c = self.counts_raw()
        return {"bases": b, "valid_kmers": v, "candidates": cd, "gap_windows": g, "contigs": c, "minimizers": self.n}

    def write_tsv(self, path, pos=True, strand=False, seq=True):
        """Write the `indexlr` text format (id \\t hash:pos:seq ...; bin/ntjoin_utils.py:173-185 reads it)."""
        lib = self._e._lib
        check(lib, lib.mxe_write_tsv(self._h, str(path).encode(), int(pos), int(strand), int(seq)))

    def per_contig(self):
        """[(name, out_hash[], pos[]) ...] for every record, in input order (empty arrays allowed)."""
        oh, ps, cg = self.out_hash, self.pos, self.contig
        bounds = np.searchsorted(cg, np.arange(len(self.names) + 1))
        return [(self.names[c], oh[bounds[c]:bounds[c + 1]], ps[bounds[c]:bounds[c + 1]]) for c in range(len(self.names))]

    def close(self):
        if self._h:
            self._e._lib.mxe_sketch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SketchSlice:
    """The minimizers of one assembly inside a sketch of several assemblies made in one call
    (Engine.sketch_device_multi): a contiguous record range [rec0, rec1) = a contiguous index range [i0, i1)."""

    def __init__(self, parent, rec0, rec1, i0, i1):
        self._p, self.rec0, self.rec1, self.i0, self.i1 = parent, int(rec0), int(rec1), int(i0), int(i1)
        self.names = parent.names[self.rec0:self.rec1]
        self._e = parent._e

    n = property(lambda s: s.i1 - s.i0)
    out_hash = property(lambda s: s._p.out_hash[s.i0:s.i1])
    min_hash = property(lambda s: s._p.min_hash[s.i0:s.i1])
    pos = property(lambda s: s._p.pos[s.i0:s.i1])
    contig = property(lambda s: s._p.contig[s.i0:s.i1] - np.uint32(s.rec0))
    forward = property(lambda s: s._p.forward[s.i0:s.i1])

    def device_pointers(self):
        n, ph, pp, pc = self._p.device_pointers()
        return (self.n, ph + 8 * self.i0, pp + 4 * self.i0, pc + 4 * self.i0)

    def fetch(self, copy=True):
        return tuple(x[self.i0:self.i1] for x in self._p.fetch(copy=copy))

    def close(self):
        pass          # the parent owns the arrays


class FilterResult:
    """Outcome of steps 2-3: per-assembly flags + the weighted edge list.  Arrays stay on the device
    until first touched (then one pinned device->host copy)."""

    def __init__(self, engine, handle, n_asm):
        self._e, self._h, self.n_asm = engine, handle, n_asm
        self._cache = None
        engine._live.add(self)

    def counts(self):
        """(total minimizers, vertices, edges) without copying the arrays to the host."""
        lib = self._e._lib
        a = [C.c_uint64() for _ in range(3)]
        check(lib, lib.mxe_result_counts(self._h, *[C.byref(x) for x in a]))
        return tuple(x.value for x in a)

    def fetch(self, copy=True):
        """Bring flags, vertices and the weighted edge list to host memory (one pinned device->host copy).
        copy=False returns views into the engine's pinned block, valid until close()."""
        if not copy:
            return self._fetch_views()
        if self._cache is None:
            lib = self._e._lib
            uniq, keep = [], []
            for a in range(self.n_asm):
                n = C.c_uint64()
                u, k = C.c_void_p(), C.c_void_p()
                check(lib, lib.mxe_result_flags(self._h, a, C.byref(n), C.byref(u), C.byref(k)))
                uniq.append(_np_view(u.value, n.value, np.uint8).astype(bool))
                keep.append(_np_view(k.value, n.value, np.uint8).astype(bool))
            nv, ne = C.c_uint64(), C.c_uint64()
            p = [C.c_void_p() for _ in range(5)]
            check(lib, lib.mxe_result_graph(self._h, C.byref(nv), C.byref(p[0]), C.byref(ne), *[C.byref(x) for x in p[1:]]))
            self._cache = {
                "uniq": uniq, "keep": keep,
                "vertices": _np_view(p[0].value, nv.value, np.uint64).copy(),
                "edge_u": _np_view(p[1].value, ne.value, np.uint64).copy(),
                "edge_v": _np_view(p[2].value, ne.value, np.uint64).copy(),
                "support": _np_view(p[3].value, ne.value, np.uint32).copy(),
                "weight": _np_view(p[4].value, ne.value, np.float64).copy(),
            }
        return self._cache

    def _fetch_views(self):
        lib = self._e._lib
        out = {"uniq": [], "keep": []}
        for a in range(self.n_asm):
            n = C.c_uint64()
            u, k = C.c_void_p(), C.c_void_p()
            check(lib, lib.mxe_result_flags(self._h, a, C.byref(n), C.byref(u), C.byref(k)))
            out["uniq"].append(_np_view(u.value, n.value, np.uint8))
            out["keep"].append(_np_view(k.value, n.value, np.uint8))
        nv, ne = C.c_uint64(), C.c_uint64()
        p = [C.c_void_p() for _ in range(5)]
        check(lib, lib.mxe_result_graph(self._h, C.byref(nv), C.byref(p[0]), C.byref(ne), *[C.byref(x) for x in p[1:]]))
        for key, ptr, n, dt in (("vertices", p[0], nv, np.uint64), ("edge_u", p[1], ne, np.uint64), ("edge_v", p[2], ne, np.uint64),
                                ("support", p[3], ne, np.uint32), ("weight", p[4], ne, np.float64)):
            out[key] = _np_view(ptr.value, n.value, dt)
        return out

    @property
    def edge_keys(self):
        """Multi-GPU shards only: the global order key of every edge (ascending inside the shard)."""
        lib = self._e._lib
        n, p = C.c_uint64(), C.c_void_p()
        check(lib, lib.mxe_result_edge_keys(self._h, C.byref(n), C.byref(p)))
        return _np_view(p.value, n.value, np.uint64).copy()

    uniq = property(lambda s: s.fetch()["uniq"])
    keep = property(lambda s: s.fetch()["keep"])
    vertices = property(lambda s: s.fetch()["vertices"])
    edge_u = property(lambda s: s.fetch()["edge_u"])
    edge_v = property(lambda s: s.fetch()["edge_v"])
    support = property(lambda s: s.fetch()["support"])
    weight = property(lambda s: s.fetch()["weight"])

    def close(self):
        if self._h:
            self._e._lib.mxe_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One engine per process and GPU (mxe_create)."""

    def __init__(self, device=0, timing=False):
        self._lib = load_library()
        h = C.c_void_p()
        check(self._lib, self._lib.mxe_create(int(device), C.byref(h)))
        self._h = h
        self._live = weakref.WeakSet()
        self.device = int(device)
        if timing:
            self.set_option("timing", 1)

    def set_stream(self, cuda_stream):
        """Issue all engine work on `cuda_stream` (int handle, e.g. torch.cuda.current_stream().cuda_stream)."""
        check(self._lib, self._lib.mxe_set_stream(self._h, C.c_void_p(int(cuda_stream) if cuda_stream else None)))

    def set_option(self, name, value):
        check(self._lib, self._lib.mxe_set_option(self._h, name.encode(), float(value)))

    @staticmethod
    def _flags(canonical):
        return CANON[canonical]

    @staticmethod
    def _names(names, n):
        if names is None:
            return None, [str(i) for i in range(n)]
        arr = (C.c_char_p * n)(*[str(x).encode() for x in names])
        return arr, [str(x) for x in names]

    def _named(self, out):
        sk = Sketch(self, out, [])
        nm = C.c_char_p()
        n_contigs = sk.counts_raw()[4]
        for i in range(n_contigs):
            check(self._lib, self._lib.mxe_sketch_contig_name(out, i, C.byref(nm)))
            sk.names.append(nm.value.decode("utf-8", "replace"))
        return sk

    def sketch_file(self, path, k, w, canonical="sum"):
        out = C.c_void_p()
        check(self._lib, self._lib.mxe_sketch_file(self._h, str(path).encode(), int(k), int(w), self._flags(canonical), C.byref(out)))
        return self._named(out)

    def load_tsv(self, path):
        """Sketch object from an existing <fasta>.k<k>.w<w>.tsv (out_hash, pos, record ids and names only)."""
        out = C.c_void_p()
        check(self._lib, self._lib.mxe_sketch_load_tsv(self._h, str(path).encode(), C.byref(out)))
        return self._named(out)

    @staticmethod
    def _host_ptr(seq):
        if hasattr(seq, "data_ptr"):
            return seq.data_ptr(), seq, seq.numel()
        a = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray, memoryview)) else np.ascontiguousarray(seq, dtype=np.uint8)
        return a.ctypes.data, a, a.size

    def prefetch(self, seq):
        """Start the host->device copy of a buffer that a later sketch_buffers(seq, ...) will use (the SAME object:
        numpy uint8 array or CPU torch tensor, ideally pinned).  Returns at once; lets the copy of the next assembly
        overlap the sketch of the previous one.  Returns the object to pass to sketch_buffers."""
        ptr, keep, n = self._host_ptr(seq)
        check(self._lib, self._lib.mxe_prefetch_buffers(self._h, C.c_void_p(ptr), n))
        # the copy is asynchronous: the engine keeps the buffer alive until it has been sketched (or the engine closes),
        # so that a caller who drops its own reference cannot have the memory recycled under the copy
        self._prefetched = getattr(self, "_prefetched", [])[-3:] + [keep]
        return keep

    def sketch_many(self, assemblies, k, w, canonical="sum", prefetch_host=False):
        """assemblies: [(seq, offsets[, names])] in assembly order.  Sketches them one after the other with the
        host->device copy of each assembly overlapping the sketch of the one before.  prefetch_host=True: the caller
        wants the minimizer tuples in host memory -- every sketch starts its device->host copy as soon as it is done,
        beside the next assembly's host->device copy (the other PCIe direction) and sketch."""
        bufs = [self._host_ptr(a[0])[1] for a in assemblies]
        out = []
        for i, a in enumerate(assemblies):
            for b in bufs[i:i + 2]:
                self.prefetch(b)
            out.append(self.sketch_buffers(bufs[i], a[1], k, w, names=a[2] if len(a) > 2 else None, canonical=canonical))
            if prefetch_host:
                out[-1].prefetch_host()
        return out

    def sketch_buffers(self, seq, offsets, k, w, names=None, canonical="sum"):
        """seq: host bytes / numpy uint8 / CPU torch uint8 tensor (pinned memory is copied fastest)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        ptr, keep, _ = self._host_ptr(seq)
        carr, pynames = self._names(names, n)
        out = C.c_void_p()
        check(self._lib, self._lib.mxe_sketch_buffers(self._h, C.c_void_p(ptr), offsets.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                     n, carr, int(k), int(w), self._flags(canonical), C.byref(out)))
        return Sketch(self, out, pynames, keepalive=keep)

    def sketch_device(self, dptr, offsets, k, w, names=None, canonical="sum"):
        """dptr: raw device address (e.g. torch_tensor.data_ptr()) of the concatenated ASCII sequence."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        carr, pynames = self._names(names, n)
        out = C.c_void_p()
        check(self._lib, self._lib.mxe_sketch_device(self._h, C.c_void_p(int(dptr)), offsets.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                    n, carr, int(k), int(w), self._flags(canonical), C.byref(out)))
        return Sketch(self, out, pynames)

    def sketch_device_many(self, dptrs, offsets_list, k, w, canonical="sum"):
        """Several device-resident assemblies, each in its own buffer: enqueued on two streams so that they run
        concurrently (mxe_sketch_device_many).  Returns one Sketch per assembly."""
        n = len(dptrs)
        offs = [np.ascontiguousarray(o, dtype=np.uint64) for o in offsets_list]
        dp = (C.c_void_p * n)(*[C.c_void_p(int(p) or None) for p in dptrs])
        op = (C.POINTER(C.c_uint64) * n)(*[o.ctypes.data_as(C.POINTER(C.c_uint64)) for o in offs])
        nc = (C.c_uint32 * n)(*[len(o) - 1 for o in offs])
        out = (C.c_void_p * n)()
        check(self._lib, self._lib.mxe_sketch_device_many(self._h, n, dp, op, nc, int(k), int(w), self._flags(canonical), out))
        return [Sketch(self, C.c_void_p(out[a]), [str(i) for i in range(len(offs[a]) - 1)]) for a in range(n)]

    def sketch_device_multi(self, dptr, offsets_list, k, w, canonical="sum", starts=None):
        """Several assemblies resident in ONE device buffer, sketched in one call: offsets_list[a] are the record starts
        of assembly a relative to the assembly's own first base (n_records + 1 entries); starts[a] = byte offset of
        assembly a in the buffer (default: back to back; a gap between two assemblies -- e.g. alignment padding --
        becomes a record of its own that belongs to no assembly).  Returns (parent sketch, [SketchSlice per assembly]).
        Windows never cross records, so every slice equals the sketch of its assembly alone; what is saved is one set
        of launches and host round trips per extra assembly, which is what a small per-GPU share of a multi-GPU job is
        made of."""
        offs, ranges, at, nrec = [np.zeros(1, dtype=np.uint64)], [], 0, 0
        for a, o in enumerate(offsets_list):
            o = np.ascontiguousarray(o, dtype=np.uint64)
            begin = at if starts is None else int(starts[a])
            if begin < at:
                raise ValueError("assemblies overlap in the buffer")
            if begin > at:                                   # gap record
                offs.append(np.array([begin], dtype=np.uint64))
                nrec += 1
            offs.append(o[1:] + np.uint64(begin))
            ranges.append((nrec, nrec + len(o) - 1))
            nrec += len(o) - 1
            at = begin + int(o[-1])
        parent = self.sketch_device(dptr, np.concatenate(offs), k, w, canonical=canonical)
        out = []
        for r0, r1 in ranges:
            idx = []
            for b in (r0, r1):
                if b == 0:
                    idx.append(0)
                elif b == nrec:
                    idx.append(parent.n)
                else:
                    v = C.c_uint64()
                    check(self._lib, self._lib.mxe_sketch_record_start(parent._h, int(b), C.byref(v)))
                    idx.append(v.value)
            out.append(SketchSlice(parent, r0, r1, idx[0], idx[1]))
        return parent, out

    def filter_and_edges(self, sketches, weights):
        """sketches in assembly order: references (CLI order) then target (bin/ntjoin.py:181-185)."""
        if any(isinstance(s, SketchSlice) for s in sketches):
            ptrs = [s.device_pointers() for s in sketches]
            return self.filter_and_edges_device([p[1] for p in ptrs], [p[3] for p in ptrs], [p[0] for p in ptrs], weights)
        n = len(sketches)
        hs = (C.c_void_p * n)(*[s._h for s in sketches])
        ws = (C.c_double * n)(*[float(x) for x in weights])
        out = C.c_void_p()
        check(self._lib, self._lib.mxe_filter_and_edges(self._h, hs, n, ws, C.byref(out)))
        return FilterResult(self, out, n)

    def filter_and_edges_device(self, d_hash, d_contig, counts, weights):
        """Raw device arrays per assembly (after a multi-GPU gather): uint64 hashes, uint32 record ids."""
        n = len(counts)
        dh = (C.c_void_p * n)(*[int(x) for x in d_hash])
        dc = (C.c_void_p * n)(*[int(x) for x in d_contig])
        cn = (C.c_uint64 * n)(*[int(x) for x in counts])
        ws = (C.c_double * n)(*[float(x) for x in weights])
        out = C.c_void_p()
        check(self._lib, self._lib.mxe_filter_and_edges_device(self._h, dh, dc, cn, n, ws, C.byref(out)))
        return FilterResult(self, out, n)

    def dist_stages(self):
        """Engine-backed stages of the multi-GPU steps 2-3, all-reduce formulation (ntjoin_b200.dist; default)."""
        return EngineDistStages(self)

    def a2a_stages(self):
        """Engine-backed stages of the multi-GPU steps 2-3, all-to-all formulation (ntjoin_b200.dist; work ~ 1/world)."""
        return EngineA2AStages(self)

    def p2p(self, rank, world, cap_total, n_asm_max=4):
        """Steps 2-3 across `world` GPUs with direct peer stores over NVLink (include/mxe.h: mxe_p2p_*); see P2PFilter."""
        return P2PFilter(self, rank, world, cap_total, n_asm_max)

    def timing(self, name):
        ms, nl = C.c_double(), C.c_uint64()
        check(self._lib, self._lib.mxe_timing(self._h, name.encode(), C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def timing_reset(self):
        check(self._lib, self._lib.mxe_timing_reset(self._h))

    def kernel_launches(self):
        return int(self._lib.mxe_kernel_launches(self._h))

    def close(self):
        if self._h:
            for sk in list(self._live):   # sketches hold engine-owned memory: release them first
                sk.close()
            self._lib.mxe_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass



class P2PFilter:
    """One rank of the peer-store formulation of steps 2-3 (csrc/p2p.cu).  Life cycle: create on every rank, exchange
    the 64-byte IPC handles (`handle()` -> `connect(handles)`; ranks that share a process use `workspace()` ->
    `connect_pointers`), then per job the five stages in order on every rank (`run` issues them back to back)."""

    STAGES = ("scatter", "buckets", "adjacency", "edges", "finish")

    def __init__(self, engine, rank, world, cap_total, n_asm_max=4):
        self._e, self._lib, self.rank, self.world = engine, engine._lib, int(rank), int(world)
        h = C.c_void_p()
        check(self._lib, self._lib.mxe_p2p_create(engine._h, int(rank), int(world), int(cap_total), int(n_asm_max), C.byref(h)))
        self._h = h
        self._keep = None

    def handle(self):
        buf = C.create_string_buffer(64)
        check(self._lib, self._lib.mxe_p2p_handle(self._h, buf, None))
        return buf.raw

    def connect(self, handles):
        """handles: the 64-byte handles of all ranks, in rank order"""
        blob = b"".join(handles)
        check(self._lib, self._lib.mxe_p2p_connect(self._h, C.c_char_p(blob)))

    def workspace(self):
        p = C.c_void_p()
        check(self._lib, self._lib.mxe_p2p_workspace(self._h, C.byref(p)))
        return p.value

    def connect_pointers(self, bases):
        arr = (C.c_void_p * len(bases))(*[C.c_void_p(int(b)) for b in bases])
        check(self._lib, self._lib.mxe_p2p_connect_pointers(self._h, arr))

    def scatter(self, hash_ptrs, contig_ptrs, counts, weights):
        """raw device addresses of this rank's out_hash (uint64) / record id (uint32) arrays per assembly"""
        n = len(counts)
        dh = (C.c_void_p * n)(*[C.c_void_p(int(p) or None) for p in hash_ptrs])
        dc = (C.c_void_p * n)(*[C.c_void_p(int(p) or None) for p in contig_ptrs])
        cn = (C.c_uint64 * n)(*[int(x) for x in counts])
        ws = (C.c_double * n)(*[float(x) for x in weights])
        self._n_asm = n
        check(self._lib, self._lib.mxe_p2p_scatter(self._h, dh, dc, cn, n, ws))

    def scatter_sketches(self, sketches, weights):
        ptrs = [sk.device_pointers() for sk in sketches]
        self._keep = list(sketches)
        self.scatter([p[1] for p in ptrs], [p[3] for p in ptrs], [p[0] for p in ptrs], weights)

    def buckets(self):
        check(self._lib, self._lib.mxe_p2p_buckets(self._h))

    def adjacency(self):
        check(self._lib, self._lib.mxe_p2p_adjacency(self._h))

    def edges(self):
        check(self._lib, self._lib.mxe_p2p_edges(self._h))

    def finish(self):
        out = C.c_void_p()
        check(self._lib, self._lib.mxe_p2p_finish(self._h, C.byref(out)))
        self._keep = None
        return FilterResult(self._e, out, self._n_asm)

    def run(self, sketches, weights):
        """all five stages for this rank (the other ranks run theirs in their own processes)"""
        self.scatter_sketches(sketches, weights)
        self.buckets()
        self.adjacency()
        self.edges()
        return self.finish()

    def close(self):
        if self._h:
            self._lib.mxe_p2p_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _u64arr(values):
    return (C.c_uint64 * len(values))(*[int(x) for x in values])


class EngineDistStages:
    """The four device stages of the multi-GPU steps 2-3 (include/mxe.h: mxe_dist_*), taking torch
    tensors that live on this engine's device.  The collectives between them are issued by
    ntjoin_b200.dist on the same tensors."""

    def __init__(self, engine):
        self._e = engine
        self._lib = engine._lib

    def mark(self, keys, asm_off, rank, world, mk):
        h, nv = C.c_void_p(), C.c_uint64()
        check(self._lib, self._lib.mxe_dist_mark(self._e._h, C.c_void_p(keys.data_ptr()), _u64arr(asm_off), len(asm_off) - 1,
                                                 int(rank), int(world), C.c_void_p(mk.data_ptr()), C.byref(h), C.byref(nv)))
        return h, nv.value

    def adjacency(self, handle, mk, vbase, loc_off, loc_n, contigs, succ):
        n = len(contigs)
        cp = (C.c_void_p * n)(*[c.data_ptr() if c.numel() else None for c in contigs])
        check(self._lib, self._lib.mxe_dist_adjacency(handle, C.c_void_p(mk.data_ptr()), _u64arr(vbase), _u64arr(loc_off), _u64arr(loc_n),
                                                      cp, C.c_void_p(succ.data_ptr())))

    def edges(self, handle, succ, srcmin):
        ne = C.c_uint64()
        check(self._lib, self._lib.mxe_dist_edges(handle, C.c_void_p(succ.data_ptr()), C.c_void_p(srcmin.data_ptr()), C.byref(ne)))
        return ne.value

    def finish(self, handle, srcmin, weights):
        n = len(weights)
        ws = (C.c_double * n)(*[float(x) for x in weights])
        out = C.c_void_p()
        try:
            check(self._lib, self._lib.mxe_dist_finish(handle, C.c_void_p(srcmin.data_ptr()), ws, C.byref(out)))
        finally:
            self._lib.mxe_dist_free(handle)
        return FilterResult(self._e, out, n)

    def abort(self, handle):
        self._lib.mxe_dist_free(handle)


class EngineA2AStages:
    """The four device stages of the all-to-all formulation (include/mxe.h: mxe_a2a_*).  Tensors live on this engine's
    device; send buffers returned by a stage are engine-owned views valid until finish()."""

    def __init__(self, engine):
        self._e = engine
        self._lib = engine._lib

    def _view(self, ptr, n, dtype):
        import torch
        from .dist import DeviceArray
        dev = torch.device("cuda", self._e.device)
        if not n:
            return torch.empty(0, dtype=dtype, device=dev)
        return torch.as_tensor(DeviceArray(ptr, n, "<i8"), device=dev)

    def partition(self, hashes, rank, world):
        import torch
        n_asm = len(hashes)
        hp = (C.c_void_p * n_asm)(*[h.data_ptr() if h.numel() else None for h in hashes])
        ns = _u64arr([h.numel() for h in hashes])
        counts = (C.c_uint64 * (world * n_asm))()
        h, send = C.c_void_p(), C.c_void_p()
        check(self._lib, self._lib.mxe_a2a_partition(self._e._h, hp, ns, n_asm, int(rank), int(world), C.byref(h), counts, C.byref(send)))
        total = sum(int(x.numel()) for x in hashes)
        cnt = np.frombuffer(counts, dtype=np.uint64).astype(np.int64).reshape(world, n_asm)
        return h, cnt, self._view(send.value, total, torch.int64)

    def mark(self, handle, recv_keys, recv_counts, ret_marks):
        nv = C.c_uint64()
        check(self._lib, self._lib.mxe_a2a_mark(handle, C.c_void_p(recv_keys.data_ptr() if recv_keys.numel() else None),
                                                _u64arr(np.asarray(recv_counts).reshape(-1)),
                                                C.c_void_p(ret_marks.data_ptr() if ret_marks.numel() else None), C.byref(nv)))
        return nv.value

    def sightings(self, handle, marks, contigs, goff, world):
        import torch
        n = len(contigs)
        cp = (C.c_void_p * n)(*[c.data_ptr() if c.numel() else None for c in contigs])
        rc = (C.c_uint64 * world)()
        send = C.c_void_p()
        check(self._lib, self._lib.mxe_a2a_sightings(handle, C.c_void_p(marks.data_ptr() if marks.numel() else None), cp, _u64arr(goff),
                                                     rc, C.byref(send)))
        cnt = np.frombuffer(rc, dtype=np.uint64).astype(np.int64)
        return cnt, self._view(send.value, 3 * int(cnt.sum()), torch.int64)

    def finish(self, handle, recv_records, n_records, n_global, weights):
        n = len(weights)
        ws = (C.c_double * n)(*[float(x) for x in weights])
        out = C.c_void_p()
        try:
            check(self._lib, self._lib.mxe_a2a_finish(handle, C.c_void_p(recv_records.data_ptr() if recv_records.numel() else None),
                                                      int(n_records), int(n_global), ws, C.byref(out)))
        finally:
            self._lib.mxe_a2a_free(handle)
        return FilterResult(self._e, out, n)

    def abort(self, handle):
        self._lib.mxe_a2a_free(handle)


class HostSketch:
    """Minimizers of one assembly held in host arrays (no engine, no GPU): mxe_sketch_from_arrays.  Used to write the
    assembly's one `.tsv` (seam S2) from the record ranges that several GPUs sketched; see dist.gather_and_write_tsv."""

    def __init__(self, out_hash, pos, contig, names, k, forward=None, min_hash=None, offsets=None, seq=None):
        self._lib = load_library()
        oh = np.ascontiguousarray(out_hash, dtype=np.uint64)
        ps = np.ascontiguousarray(pos, dtype=np.uint32)
        cg = np.ascontiguousarray(contig, dtype=np.uint32)
        fw = None if forward is None else np.ascontiguousarray(forward, dtype=np.uint8)
        mh = None if min_hash is None else np.ascontiguousarray(min_hash, dtype=np.uint64)
        of = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.uint64)
        self._seq = None if seq is None else (np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray, memoryview))
                                              else np.ascontiguousarray(seq, dtype=np.uint8))     # borrowed by the C side: keep alive
        nm = (C.c_char_p * max(1, len(names)))(*[str(x).encode() for x in names])
        h = C.c_void_p()
        ptr = lambda a: None if a is None else C.c_void_p(a.ctypes.data)      # noqa: E731
        check(self._lib, self._lib.mxe_sketch_from_arrays(ptr(oh), ptr(mh), ptr(ps), ptr(cg), ptr(fw), len(oh), nm,
                                                          None if of is None else of.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                          len(names), int(k), ptr(self._seq), C.byref(h)))
        self._h, self.n, self.names = h, len(oh), [str(x) for x in names]

    def write_tsv(self, path, pos=True, strand=False, seq=True):
        check(self._lib, self._lib.mxe_write_tsv(self._h, str(path).encode(), int(pos), int(strand), int(seq)))

    def close(self):
        if self._h:
            self._lib.mxe_sketch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
