"""`<prefix>.mx.dot` from arrays.

The reference writes the minimizer graph with a Python loop over igraph vertex and edge objects
(Ntjoin.print_graph, bin/ntjoin.py:25-67): one f-string per vertex with a dictionary lookup per assembly, one
name lookup per edge endpoint.  `write_mx_dot` produces the same bytes from the arrays the engine already holds;
the formatting that is Python's own (repr of the record names in the label tuples, str of the float weights)
is done here once per distinct value, the per-vertex / per-edge text in C (mxe_write_dot, host only).
"""
import ctypes as C

import numpy as np

from ._lib import check, load_library

COLOURS = ["red", "green", "blue", "purple", "orange", "turquoise", "pink", "yellow", "orchid", "salmon"]   # bin/ntjoin.py:38-39


def edge_attr_texts(masks, asm_keys, weights):
    """distinct support masks -> ' [weight=<float> color=<colour>]\\n' (bin/ntjoin.py:52-60, bin/ntjoin_utils.py:54-56)"""
    n_asm = len(asm_keys)
    colours = COLOURS if n_asm <= len(COLOURS) else ["red"] * n_asm
    uniq, inverse = np.unique(np.asarray(masks, dtype=np.uint32), return_inverse=True)
    texts = []
    for m in uniq.tolist():
        support = [a for a in range(n_asm) if m >> a & 1]
        weight = sum(weights[asm_keys[a]] if isinstance(weights, dict) else weights[a] for a in support)   # Python's sum(): starts at int 0
        colour = colours[support[0]] if len(support) == 1 else "lightgrey" if len(support) == 2 else "black"
        texts.append(f" [weight={weight} color={colour}]\n")
    return texts, inverse.astype(np.uint32)


def write_mx_dot(path, vertices, asm_keys, record_names, v_ctg, v_pos, e_src, e_dst, support_mask, weights):
    """vertices: uint64 hashes in graph.vs order; record_names[a]: names of assembly a's records; v_ctg[a] / v_pos[a]:
    record index and position of every vertex in assembly a; e_src / e_dst: vertex indices as igraph reports
    edge.source / edge.target; support_mask: bit a = assembly a; weights: per assembly (list) or by key (dict)."""
    lib = load_library()
    n_asm = len(asm_keys)
    vertices = np.ascontiguousarray(vertices, dtype=np.uint64)
    n_v = len(vertices)
    keys = (C.c_char_p * n_asm)(*[str(k).encode() for k in asm_keys])
    repr_arrays = [(C.c_char_p * max(1, len(names)))(*[repr(str(x)).encode() for x in names]) for names in record_names]
    reprs = (C.c_void_p * n_asm)(*[C.cast(r, C.c_void_p) for r in repr_arrays])
    ctg = [np.ascontiguousarray(x, dtype=np.uint32) for x in v_ctg]
    pos = [np.ascontiguousarray(x, dtype=np.uint32) for x in v_pos]
    assert all(len(x) == n_v for x in ctg + pos)
    ctg_p = (C.c_void_p * n_asm)(*[x.ctypes.data for x in ctg])
    pos_p = (C.c_void_p * n_asm)(*[x.ctypes.data for x in pos])
    e_src = np.ascontiguousarray(e_src, dtype=np.uint32)
    e_dst = np.ascontiguousarray(e_dst, dtype=np.uint32)
    texts, attr = edge_attr_texts(support_mask, list(asm_keys), weights)
    attr = np.ascontiguousarray(attr, dtype=np.uint32)
    text_p = (C.c_char_p * max(1, len(texts)))(*[t.encode() for t in texts])
    check(lib, lib.mxe_write_dot(str(path).encode(), n_v, C.c_void_p(vertices.ctypes.data), n_asm, keys, reprs, ctg_p, pos_p,
                                 len(e_src), C.c_void_p(e_src.ctypes.data), C.c_void_p(e_dst.ctypes.data),
                                 C.c_void_p(attr.ctypes.data), text_p))
