"""Engine-backed replacements for ntJoin's three step-2/3 Python functions (seam S3).

Reference signatures (bin/ntjoin_utils.py):
    read_minimizers(tsv_filename, repeat_bf=False) -> (mx_info: dict[str,(str,int)], mxs: list[list[str]])   :167-193
    filter_minimizers(list_mxs: dict[asm, list[list[str]]]) -> same shape                                     :152-165
    build_graph(list_mxs, weights, graph=None, black_list=None) -> igraph.Graph (vs['name'], es['support'],
                                                                               es['weight'])                 :83-141

`install(ntjoin_utils_module)` swaps them in; `install_print_graph(ntjoin_module)` does the same for
Ntjoin.print_graph (bin/ntjoin.py:25-67), which then writes `<prefix>.mx.dot` from the engine's arrays.  Return types and contents are identical to the
reference's; the arithmetic (uniqueness counting, intersection, adjacent-pair edge reduction, weights)
runs on the GPU through the C ABI.  The lists returned by read_minimizers / filter_minimizers carry a
hidden handle to the device-resident arrays so the next stage does not re-parse strings.  Calls the
engine cannot serve (repeat_bf, black_list, an existing graph, plain lists from other callers such as
bin/ntjoin_overlap.py:25-28) fall through to the reference's original functions.
"""
import contextlib
import gc

import numpy as np

_ENGINE = None


@contextlib.contextmanager
def _gc_paused():
    """The functions below create millions of small acyclic objects (strings, tuples, per-record lists) in one go; every
    few hundred thousand container allocations the cyclic collector walks all of them again and finds nothing.  Pausing
    it for the duration of the bulk creation takes 40 % off build_graph (0.83 -> 0.51 s for 375 k edges)."""
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


def _trace(what):
    """MXE_DROPIN_TRACE=<file>: one line per seam call, saying whether the engine served it or the original ran"""
    import os
    path = os.environ.get("MXE_DROPIN_TRACE")
    if path:
        with open(path, "a", encoding="utf-8") as fh:
            fh.write(what + "\n")


def _shutdown():
    """release the cached result and the engine in order (the interpreter frees module globals in any order)"""
    global _ENGINE, _LAST_FILTER
    if _LAST_FILTER is not None:
        try:
            _LAST_FILTER[1].close()
        except Exception:
            pass
        _LAST_FILTER = None
    if _ENGINE is not None:
        try:
            _ENGINE.close()
        except Exception:
            pass
        _ENGINE = None


def _engine():
    global _ENGINE
    if _ENGINE is None:
        import atexit
        import os
        from .engine import Engine
        _ENGINE = Engine(int(os.environ.get("MXE_DEVICE", "0")))
        atexit.register(_shutdown)
    return _ENGINE


class MxLists(list):
    """list[list[str]] plus the arrays it was built from."""
    _sketch = None      # Sketch (device-resident out_hash / contig)
    _mask = None        # bool mask over the sketch selecting the entries present in these lists
    _result = None      # FilterResult shared by all assemblies after filter_minimizers
    _asm_index = None
    _strs = None        # decimal strings of every minimizer of the sketch (made once, in read_minimizers)

    def __reduce__(self):
        # The reference keeps these lists in Ntjoin.list_mxs and pickles `self` into multiprocessing.Pool workers when
        # -t / assemble_t > 1 (bin/ntjoin.py:173-174).  The engine handles (ctypes pointers) cannot and need not travel:
        # the lists pickle as the plain list[list[str]] the reference expects.
        return (list, (list(self),))


MAX_ASSEMBLIES = 32              # support masks are 32 bits wide (include/mxe.h: n_asm <= 32)
MAX_MINIMIZERS = (1 << 32) - 1   # steps 2-3 index minimizers with 32 bits


def _engine_can_serve(vals):
    return len(vals) <= MAX_ASSEMBLIES and sum(v._sketch.n for v in vals) < MAX_MINIMIZERS


_LAST_FILTER = None              # (tuple of sketch ids, FilterResult): filter_minimizers and build_graph share one pass


def _filter_once(eng, vals):
    """steps 2-3 of these sketches, computed once per set of assemblies: uniqueness, intersection and the edge list do
    not depend on the weights (a weight is the sum over the support mask, re-derived in build_graph)"""
    global _LAST_FILTER
    key = tuple(id(v._sketch) for v in vals)
    if _LAST_FILTER is not None and _LAST_FILTER[0] == key:
        _trace("filter_and_edges cached")
        return _LAST_FILTER[1]
    if _LAST_FILTER is not None:
        _LAST_FILTER[1].close()
    _trace("filter_and_edges engine")
    res = eng.filter_and_edges([v._sketch for v in vals], [1.0] * len(vals))
    _LAST_FILTER = (key, res, [v._sketch for v in vals])      # the sketches stay referenced: ids cannot be recycled
    return res


def _lists_from(sk, mask, strs=None):
    """per-record lists (records with at least one minimizer in the TSV) of decimal strings"""
    oh, cg = sk.out_hash, sk.contig
    if strs is None:
        strs = np.array([str(h) for h in oh.tolist()], dtype=object)
    present = np.unique(cg)
    bounds = np.searchsorted(cg, np.arange(len(sk.names) + 1))
    out = MxLists()
    for c in present:
        s, e = bounds[c], bounds[c + 1]
        out.append(strs[s:e][mask[s:e]].tolist())
    return out, strs


def make_read_minimizers(original):
    def read_minimizers(tsv_filename, repeat_bf=False):
        if repeat_bf:
            _trace("read_minimizers original")
            return original(tsv_filename, repeat_bf)
        _trace("read_minimizers engine")
        eng = _engine()
        sk = eng.load_tsv(tsv_filename)
        res = eng.filter_and_edges([sk], [1.0])
        uniq = res.uniq[0]
        with _gc_paused():
            mxs, strs = _lists_from(sk, uniq)
            names = sk.names
            cg, ps = sk.contig[uniq].tolist(), sk.pos[uniq].tolist()
            mx_info = {h: (names[c], p) for h, c, p in zip(strs[uniq].tolist(), cg, ps)}
        mxs._sketch, mxs._mask, mxs._strs = sk, uniq, strs
        res.close()
        return mx_info, mxs
    read_minimizers.__doc__ = original.__doc__
    return read_minimizers


def make_filter_minimizers(original):
    def filter_minimizers(list_mxs):
        vals = list(list_mxs.values())
        if not vals or not all(isinstance(v, MxLists) and v._sketch is not None for v in vals) or not _engine_can_serve(vals):
            _trace("filter_minimizers original")
            return original(list_mxs)
        _trace("filter_minimizers engine")
        eng = _engine()
        res = _filter_once(eng, vals)
        out = {}
        with _gc_paused():
            for a, (asm, v) in enumerate(list_mxs.items()):
                keep = res.keep[a]
                lists, _ = _lists_from(v._sketch, keep, v._strs)
                lists._sketch, lists._mask, lists._asm_index, lists._strs = v._sketch, keep, a, v._strs
                out[asm] = lists
        return out
    filter_minimizers.__doc__ = original.__doc__
    return filter_minimizers


def make_build_graph(original, ig):
    def build_graph(list_mxs, weights, graph=None, black_list=None):
        vals = list(list_mxs.values())
        if graph is not None or black_list is not None or not vals or \
                not all(isinstance(v, MxLists) and v._sketch is not None and v._asm_index is not None for v in vals) or \
                not _engine_can_serve(vals):
            _trace("build_graph original")
            return original(list_mxs, weights, graph, black_list)
        _trace("build_graph engine")
        eng = _engine()
        keys = list(list_mxs.keys())
        res = _filter_once(eng, vals)
        g = ig.Graph()
        vs = res.vertices
        with _gc_paused():
            g.add_vertices([str(v) for v in vs.tolist()])
            eu = np.searchsorted(vs, res.edge_u)
            ev = np.searchsorted(vs, res.edge_v)
            g.add_edges(list(zip(eu.tolist(), ev.tolist())))
            n_asm = len(keys)
            masks = res.support.tolist()
            support_of = {m: [keys[a] for a in range(n_asm) if m >> a & 1] for m in set(masks)}
            # calc_total_weight (bin/ntjoin_utils.py:54-56): Python's sum() over the support list, in assembly order
            weight_of = {m: sum(weights[f] for f in sup) for m, sup in support_of.items()}
            g.es["support"] = [list(support_of[m]) for m in masks]
            g.es["weight"] = [weight_of[m] for m in masks]
            _attach_dot_payload(g, keys, vals, weights, vs, eu, ev, res)
        return g
    build_graph.__doc__ = original.__doc__
    return build_graph


class _DotPayload:
    """arrays behind a graph built by the engine-backed build_graph: lets print_graph write `.mx.dot` without
    walking igraph objects"""
    pass


def _attach_dot_payload(g, keys, vals, weights, vs, eu, ev, res):
    p = _DotPayload()
    p.keys, p.vertices, p.weights = list(keys), vs, [weights[k] for k in keys]
    p.e_src, p.e_dst = np.minimum(eu, ev).astype(np.uint32), np.maximum(eu, ev).astype(np.uint32)   # igraph: source = lower id
    p.masks = res.support.copy()
    p.names, p.v_ctg, p.v_pos = [], [], []
    for a, v in enumerate(vals):
        sk, keep = v._sketch, res.keep[a]
        order = np.argsort(sk.out_hash[keep], kind="stable")          # survivors of assembly a in vertex (ascending hash) order
        p.names.append(list(sk.names))
        p.v_ctg.append(sk.contig[keep][order])
        p.v_pos.append(sk.pos[keep][order])
    try:
        g["_mxe_dot"] = p            # python-igraph: graph attribute
    except TypeError:
        g._mxe_dot = p               # stand-ins without graph attributes


def _dot_payload(graph):
    try:
        return graph["_mxe_dot"]
    except (KeyError, TypeError, IndexError):
        return getattr(graph, "_mxe_dot", None)


def make_print_graph(original):
    """Ntjoin.print_graph (bin/ntjoin.py:25-67) with the per-vertex / per-edge text written from arrays."""
    def print_graph(self, graph, out_prefix=None):
        import datetime
        import sys
        from .dot import COLOURS, write_mx_dot
        p = _dot_payload(graph)
        count = lambda seq: seq() if callable(seq) else seq      # noqa: E731
        if p is None or list(self.list_mx_info.keys()) != p.keys or \
                len(count(graph.vs)) != len(p.vertices) or len(count(graph.es)) != len(p.e_src):
            _trace("print_graph original")
            return original(self, graph, out_prefix)
        _trace("print_graph engine")
        out_graph = self.args.p + ".mx.dot" if out_prefix is None else out_prefix + "mx.dot"
        print(datetime.datetime.today(), ": Printing graph", out_graph, sep=" ", file=sys.stdout)
        write_mx_dot(out_graph, p.vertices, p.keys, p.names, p.v_ctg, p.v_pos, p.e_src, p.e_dst, p.masks, p.weights)
        colours = COLOURS if len(p.keys) <= len(COLOURS) else ["red"] * len(p.keys)
        print("\nfile_name\tnumber\tcolour")
        for i, filename in enumerate(p.keys):
            print(filename, i, colours[i], sep="\t")
        print("", flush=True)
    print_graph.__doc__ = original.__doc__
    return print_graph


def make_find_mx_min_max(original):
    """NtjoinScaffolder.find_mx_min_max (bin/ntjoin_assemble.py:688-702): per record of the target, the smallest and
    largest position of a minimizer that is a graph vertex -- from the arrays instead of one igraph name lookup per
    unique minimizer of the target."""
    def find_mx_min_max(self, target):
        p = _dot_payload(self.graph) if self.graph is not None else None
        count = lambda seq: seq() if callable(seq) else seq      # noqa: E731
        if p is None or target not in p.keys or len(count(self.graph.vs)) != len(p.vertices):
            _trace("find_mx_min_max original")
            return original(self, target)
        _trace("find_mx_min_max engine")
        a = p.keys.index(target)
        ctg, pos = np.asarray(p.v_ctg[a], dtype=np.int64), np.asarray(p.v_pos[a], dtype=np.int64)
        if not len(ctg):
            return {}
        order = np.argsort(ctg, kind="stable")
        ctg, pos = ctg[order], pos[order]
        starts = np.flatnonzero(np.concatenate([[True], ctg[1:] != ctg[:-1]]))
        lo, hi = np.minimum.reduceat(pos, starts), np.maximum.reduceat(pos, starts)
        names = p.names[a]
        # the reference's dict is filled in (record, position) order of the target's minimizers: records ascending
        return {names[int(c)]: (int(l), int(h)) for c, l, h in zip(ctg[starts], lo, hi)}
    find_mx_min_max.__doc__ = original.__doc__
    return find_mx_min_max


def install_scaffolder(module):
    """Patch a loaded `ntjoin_assemble` module (bin/ntjoin_assemble.py) in place (idempotent)."""
    cls = getattr(module, "NtjoinScaffolder", None)
    if cls is None or getattr(cls, "_mxe_min_max", False):
        return module
    cls.find_mx_min_max = make_find_mx_min_max(cls.find_mx_min_max)
    cls._mxe_min_max = True
    return module


def install_print_graph(module):
    """Patch a loaded `ntjoin` module (bin/ntjoin.py) in place (idempotent)."""
    cls = getattr(module, "Ntjoin", None)
    if cls is None or getattr(cls, "_mxe_print_graph", False):
        return module
    cls.print_graph = make_print_graph(cls.print_graph)
    cls._mxe_print_graph = True
    return module


def install(module):
    """Patch a loaded `ntjoin_utils` module in place (idempotent)."""
    if getattr(module, "_mxe_installed", False):
        return module
    module.read_minimizers = make_read_minimizers(module.read_minimizers)
    module.filter_minimizers = make_filter_minimizers(module.filter_minimizers)
    module.build_graph = make_build_graph(module.build_graph, module.ig)
    module._mxe_installed = True
    return module
