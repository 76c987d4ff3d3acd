"""CPU test of the `.mx.dot` writer (mxe_write_dot through ntjoin_b200.dot; host only, no GPU): byte-equal to what the
reference's own Ntjoin.print_graph (bin/ntjoin.py:25-67) wrote for the same graphs (tests/golden/make_golden.py)."""
import glob
import json
import os

import numpy as np
import pytest

from ntjoin_b200.dot import edge_attr_texts, write_mx_dot


def _arrays(g):
    files, k, w = g["files"], g["k"], g["w"]
    keys = [f"{i}.{f}.k{k}.w{w}.tsv" for i, f in enumerate(files)]
    vertices = np.array([int(v) for v in g["vertices"]], dtype=np.uint64)
    index = {v: i for i, v in enumerate(g["vertices"])}
    names, v_ctg, v_pos = [], [], []
    for a in range(len(files)):
        info = g["read_minimizers"][a]["mx_info"]
        nm = sorted({c for c, _ in info.values()})
        pos_of = {c: i for i, c in enumerate(nm)}
        names.append(nm)
        v_ctg.append([pos_of[info[v][0]] for v in g["vertices"]])
        v_pos.append([info[v][1] for v in g["vertices"]])
    pairs = [sorted((index[s], index[t])) for s, t in g["edges"]]
    masks = [sum(1 << a for a in sup) for sup in g["support"]]
    return keys, vertices, names, v_ctg, v_pos, [p[0] for p in pairs], [p[1] for p in pairs], masks


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "steps23_*.json"))))
def test_dot_bytes_match_reference_print_graph(path, tmp_path, golden_dir):
    g = json.load(open(path))
    name = os.path.basename(path)[len("steps23_"):-len(".json")]
    keys, vertices, names, v_ctg, v_pos, src, dst, masks = _arrays(g)
    out = tmp_path / "g.mx.dot"
    write_mx_dot(out, vertices, keys, names, v_ctg, v_pos, src, dst, masks, g["weights"])
    want = open(os.path.join(golden_dir, "expected", f"print_graph_{name}.mx.dot"), "rb").read()
    assert out.read_bytes() == want


def test_edge_attribute_texts():
    """weights are Python floats summed in assembly order from int 0; colours as bin/ntjoin.py:52-59"""
    texts, idx = edge_attr_texts([1, 2, 3, 7, 4, 3], ["a", "b", "c"], [2.0, 2.0, 1.0])
    got = [texts[i] for i in idx]
    assert got == [" [weight=2.0 color=red]\n", " [weight=2.0 color=green]\n", " [weight=4.0 color=lightgrey]\n",
                   " [weight=5.0 color=black]\n", " [weight=1.0 color=blue]\n", " [weight=4.0 color=lightgrey]\n"]
    texts, idx = edge_attr_texts([1], [str(i) for i in range(11)], {str(i): 0.1 * (i + 1) for i in range(11)})
    assert texts == [" [weight=0.1 color=red]\n"]        # more than ten assemblies: every single-support colour is red


def test_record_names_use_python_repr(tmp_path):
    out = tmp_path / "q.mx.dot"
    write_mx_dot(out, [5, 9], ["x.tsv"], [["it's", 'plain']], [[0, 1]], [[7, 8]], [0], [1], [1], {"x.tsv": 1.5})
    assert out.read_text() == 'graph G {\n"5" [label="5\nx.tsv_("it\'s", 7)"]\n"9" [label="9\nx.tsv_(\'plain\', 8)"]\n' \
                              '"5" --"9" [weight=1.5 color=red]\n}\n'
