"""GPU: the two drop-in seams -- the `indexlr` executable (S1/S2) and the patched ntjoin_utils
functions (S3) -- against the oracle CLI and the reference-generated golden vectors."""
import glob
import json
import os
import subprocess
import sys

import pytest

import oracle_lib
import ref_py

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INDEXLR = os.path.join(ROOT, "bin", "indexlr")


def test_indexlr_cli_both_argv_forms(golden_dir, tmp_path):
    fa = os.path.join(golden_dir, "inputs", "scaf.more_seqs.fa")
    want = subprocess.check_output([oracle_lib.CLI, "--seq", "--long", "--pos", "-k", "32", "-w", "500", fa])
    got = subprocess.check_output([sys.executable, INDEXLR, "--seq", "--long", "--pos", "-k", "32", "-w", "500", "-t", "4", fa])   # ntJoin:205
    assert got == want
    out = tmp_path / "o.tsv"
    subprocess.check_call([sys.executable, INDEXLR, fa, "--seq", "--long", "--pos", "-k32", "-w500", "-t4", "-o", str(out)])     # ntjoin_utils.py:198
    assert out.read_bytes() == want


def test_indexlr_cli_fails_nonzero(tmp_path):
    r = subprocess.run([sys.executable, INDEXLR, "-k", "32", "-w", "10", str(tmp_path / "missing.fa")], capture_output=True)
    assert r.returncode != 0 and b"cannot open" in r.stderr
    assert subprocess.run([sys.executable, INDEXLR, "-k", "32"], capture_output=True).returncode != 0


def test_dropin_functions_vs_golden(golden_dir, tmp_path):
    from ntjoin_b200 import dropin
    mod = dropin.install(ref_py.as_module())
    for path in sorted(glob.glob(os.path.join(golden_dir, "steps23_*.json"))):
        g = json.load(open(path))
        list_mxs, weights = {}, {}
        for i, f in enumerate(g["files"]):
            tsv = str(tmp_path / f"{i}.{f}.k{g['k']}.w{g['w']}.tsv")
            subprocess.check_call([sys.executable, INDEXLR, "--seq", "--long", "--pos", "-k", str(g["k"]), "-w", str(g["w"]),
                                   os.path.join(golden_dir, "inputs", f), "-o", tsv])
            info, mxs = mod.read_minimizers(tsv)
            assert type(info) is dict and isinstance(mxs, list)
            assert {k: list(v) for k, v in info.items()} == g["read_minimizers"][i]["mx_info"], path
            assert list(info) == g["read_minimizers"][i]["mx_order"], path
            assert mxs == g["read_minimizers"][i]["mxs"], path
            assert all(isinstance(v, tuple) and isinstance(v[0], str) and isinstance(v[1], int) for v in info.values())
            list_mxs[tsv], weights[tsv] = mxs, g["weights"][i]
        filt = mod.filter_minimizers(list_mxs)
        assert list(filt) == list(list_mxs) and [filt[t] for t in list_mxs] == g["filter_minimizers"], path
        gr = mod.build_graph(filt, weights)
        keys = list(list_mxs)
        assert sorted(gr.vnames, key=int) == g["vertices"], path
        assert [list(e) for e in gr.edges] == g["edges"], path
        assert [[keys.index(f) for f in s] for s in gr.eattr["support"]] == g["support"], path
        assert gr.eattr["weight"] == g["weight"], path


def test_dropin_print_graph_bytes(golden_dir, tmp_path, monkeypatch):
    """Ntjoin.print_graph replaced by the array writer: same bytes as the reference's own print_graph wrote for the
    same graph (tests/golden/expected/print_graph_*.mx.dot), and the original runs when the graph is not ours"""
    import types
    from ntjoin_b200 import dropin
    mod = dropin.install(ref_py.as_module())
    calls = []

    class Ntjoin:                                        # shaped like bin/ntjoin.py's class: list_mx_info, args.p
        def print_graph(self, graph, out_prefix=None):
            calls.append(graph)

    nt = dropin.install_print_graph(types.SimpleNamespace(Ntjoin=Ntjoin))
    monkeypatch.chdir(tmp_path)
    for name in ("config1_ff_w500", "three_way_w1000", "misassembled_frrf_w500", "selfdup_w250"):
        g = json.load(open(os.path.join(golden_dir, f"steps23_{name}.json")))
        obj = nt.Ntjoin()
        obj.list_mx_info, list_mxs, weights = {}, {}, {}
        for i, f in enumerate(g["files"]):
            tsv = f"{i}.{f}.k{g['k']}.w{g['w']}.tsv"     # relative names = the assembly keys of the golden file
            subprocess.check_call([sys.executable, INDEXLR, "--seq", "--long", "--pos", "-k", str(g["k"]), "-w", str(g["w"]),
                                   os.path.join(golden_dir, "inputs", f), "-o", tsv])
            obj.list_mx_info[tsv], list_mxs[tsv] = mod.read_minimizers(tsv)
            weights[tsv] = g["weights"][i]
        gr = mod.build_graph(mod.filter_minimizers(list_mxs), weights)
        obj.args = types.SimpleNamespace(p=f"out_{name}")
        obj.print_graph(gr)
        want = open(os.path.join(golden_dir, "expected", f"print_graph_{name}.mx.dot"), "rb").read()
        assert open(f"out_{name}.mx.dot", "rb").read() == want, name
    assert not calls
    obj.print_graph(types.SimpleNamespace(vs=[], es=[]))     # not built by the engine: the original must run
    assert len(calls) == 1


def test_dropin_find_mx_min_max(golden_dir, tmp_path, monkeypatch):
    """NtjoinScaffolder.find_mx_min_max from arrays == the reference loop (bin/ntjoin_assemble.py:688-702), incl. dict order"""
    import types
    from ntjoin_b200 import dropin
    mod = dropin.install(ref_py.as_module())

    class NtjoinScaffolder:
        def find_mx_min_max(self, target):               # the reference's loop, vertex lookup by name
            vertices = {v["name"] for v in self.graph.vs}
            out = {}
            for mx, (ctg, pos) in self.list_mx_info[target].items():
                if mx in vertices:
                    out[ctg] = (min(out[ctg][0], pos), max(out[ctg][1], pos)) if ctg in out else (pos, pos)
            return out

    original = NtjoinScaffolder.find_mx_min_max
    sc = dropin.install_scaffolder(types.SimpleNamespace(NtjoinScaffolder=NtjoinScaffolder)).NtjoinScaffolder
    monkeypatch.chdir(tmp_path)
    for name in ("multiple_w500", "misassembled_ffrr_w500", "selfdup_w250", "overlap_k15_w10"):
        g = json.load(open(os.path.join(golden_dir, f"steps23_{name}.json")))
        obj = sc()
        obj.list_mx_info, list_mxs, weights = {}, {}, {}
        for i, f in enumerate(g["files"]):
            tsv = f"{i}.{f}.k{g['k']}.w{g['w']}.tsv"
            subprocess.check_call([sys.executable, INDEXLR, "--seq", "--long", "--pos", "-k", str(g["k"]), "-w", str(g["w"]),
                                   os.path.join(golden_dir, "inputs", f), "-o", tsv])
            obj.list_mx_info[tsv], list_mxs[tsv] = mod.read_minimizers(tsv)
            weights[tsv] = g["weights"][i]
        obj.graph = mod.build_graph(mod.filter_minimizers(list_mxs), weights)
        for target in list_mxs:
            got, want = obj.find_mx_min_max(target), original(obj, target)
            assert got == want and list(got) == list(want), (name, target)
            assert all(type(v[0]) is int and type(v[1]) is int for v in got.values())
        assert len(want) > 0


def test_dropin_falls_through_for_plain_lists():
    """callers like bin/ntjoin_overlap.py:25-28 pass plain lists keyed by ints: the originals must run"""
    from ntjoin_b200 import dropin
    mod = dropin.install(ref_py.as_module())
    lm = {0: [["1", "2", "3"]], 1: [["3", "2", "9"]]}
    f = mod.filter_minimizers(lm)
    assert f == {0: [["2", "3"]], 1: [["3", "2"]]}
    g = mod.build_graph(f, {0: 1, 1: 1})
    assert g.edges == [("2", "3")] and g.eattr["weight"] == [2]


def test_btllib_compat_overlap_resketch(golden_dir, oracle):
    """bin/ntjoin_assemble.py:478-481 shape: Indexlr(path, k=15, w=10, LONG_MODE, threads) -> .id, .minimizers[].out_hash/.pos"""
    import numpy as np
    from ntjoin_b200 import btllib_compat as btllib
    fa = os.path.join(golden_dir, "inputs", "scaf.f-f.overlapping.fa")
    names, seq, offs = oracle_lib.read_fasta(fa)
    ref = oracle.sketch(seq, offs, 15, 10)
    got = []
    with btllib.Indexlr(fa, 15, 10, btllib.IndexlrFlag.LONG_MODE, 4) as ix:
        for rec in ix:
            assert rec.id == names[rec.num]
            got += [(rec.num, m.out_hash, m.pos) for m in rec.minimizers]
    assert got == [(int(c), int(h), int(p)) for c, h, p in zip(ref["contig"], ref["out_hash"], ref["pos"])]
    with btllib.SeqReader(fa, btllib.SeqReaderFlag.LONG_MODE, 4) as rd:
        recs = list(rd)
    assert [r.id for r in recs] == names and "".join(r.seq for r in recs).encode() == seq
    with pytest.raises(FileNotFoundError):
        btllib.SeqReader(fa + ".missing", btllib.SeqReaderFlag.LONG_MODE, 1)
