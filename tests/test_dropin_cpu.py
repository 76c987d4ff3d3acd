"""CPU: the drop-in functions (ntjoin_b200/dropin.py: read_minimizers / filter_minimizers / build_graph) against the
fifteen step-2/3 golden cases produced by the reference's own functions, with the two engine calls they make served by
the oracle-backed stand-in of tests/harness/fake_engine.py (the GPU twin of this test, tests/test_gpu_dropin.py, runs
the same assertions with the real engine).  Checks the Python side of seam S3: types, dict order, list contents, the
graph handed to igraph."""
import glob
import json
import os
import subprocess
import sys

import oracle_lib
import ref_py

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))


def test_dropin_functions_vs_golden_with_oracle_engine(golden_dir, tmp_path, monkeypatch):
    from fake_engine import FakeEngine
    from ntjoin_b200 import dropin
    oracle_lib.build()
    monkeypatch.setattr(dropin, "_ENGINE", FakeEngine())
    monkeypatch.setattr(dropin, "_LAST_FILTER", None)
    mod = dropin.install(ref_py.as_module())
    n_cases = 0
    for path in sorted(glob.glob(os.path.join(golden_dir, "steps23_*.json"))):
        g = json.load(open(path))
        list_mxs, weights = {}, {}
        for i, f in enumerate(g["files"]):
            tsv = str(tmp_path / f"{n_cases}.{i}.{f}.k{g['k']}.w{g['w']}.tsv")
            subprocess.check_call([oracle_lib.CLI, "--seq", "--long", "--pos", "-k", str(g["k"]), "-w", str(g["w"]),
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
        assert len(gr.es) == len(g["edges"]) and len(gr.vs) == len(g["vertices"])
        if g["edges"]:
            s, t = g["edges"][0]
            assert gr.get_eid(s, t) == 0 and gr.get_eid(t, s) == 0
        n_cases += 1
    assert n_cases >= 15
    monkeypatch.setattr(dropin, "_LAST_FILTER", None)      # drop the cached result before the stand-in engine goes away


def test_dropin_gc_state_is_restored():
    """the drop-in functions pause the cyclic collector while they create their objects and leave it as they found it"""
    import gc
    from ntjoin_b200 import dropin
    assert gc.isenabled()
    with dropin._gc_paused():
        assert not gc.isenabled()
    assert gc.isenabled()
    gc.disable()
    try:
        with dropin._gc_paused():
            pass
        assert not gc.isenabled()
    finally:
        gc.enable()


def _tsvs(g, golden_dir, relative=True):
    out = []
    for i, f in enumerate(g["files"]):
        tsv = f"{i}.{f}.k{g['k']}.w{g['w']}.tsv"
        subprocess.check_call([oracle_lib.CLI, "--seq", "--long", "--pos", "-k", str(g["k"]), "-w", str(g["w"]),
                               os.path.join(golden_dir, "inputs", f), "-o", tsv])
        out.append(tsv)
    return out


def test_dropin_print_graph_and_min_max_with_oracle_engine(golden_dir, tmp_path, monkeypatch):
    """CPU twins of test_gpu_dropin.py::test_dropin_print_graph_bytes / test_dropin_find_mx_min_max: the array-fed
    `.mx.dot` writer gives the bytes of the reference's own print_graph, find_mx_min_max from arrays equals the reference's
    loop (bin/ntjoin_assemble.py:688-702) including the dict order"""
    import types
    from fake_engine import FakeEngine
    from ntjoin_b200 import dropin
    oracle_lib.build()
    monkeypatch.setattr(dropin, "_ENGINE", FakeEngine())
    monkeypatch.setattr(dropin, "_LAST_FILTER", None)
    mod = dropin.install(ref_py.as_module())
    calls = []

    class Ntjoin:
        def print_graph(self, graph, out_prefix=None):
            calls.append(graph)

    class NtjoinScaffolder:
        def find_mx_min_max(self, target):
            vertices = {v["name"] for v in self.graph.vs}
            out = {}
            for mx, (ctg, pos) in self.list_mx_info[target].items():
                if mx in vertices:
                    out[ctg] = (min(out[ctg][0], pos), max(out[ctg][1], pos)) if ctg in out else (pos, pos)
            return out

    original = NtjoinScaffolder.find_mx_min_max
    nt = dropin.install_print_graph(types.SimpleNamespace(Ntjoin=Ntjoin))
    sc = dropin.install_scaffolder(types.SimpleNamespace(NtjoinScaffolder=NtjoinScaffolder)).NtjoinScaffolder
    monkeypatch.chdir(tmp_path)
    for name in ("config1_ff_w500", "three_way_w1000", "misassembled_frrf_w500", "selfdup_w250", "multiple_w500", "overlap_k15_w10"):
        g = json.load(open(os.path.join(golden_dir, f"steps23_{name}.json")))
        obj, scaf = nt.Ntjoin(), sc()
        obj.list_mx_info, list_mxs, weights = {}, {}, {}
        for i, tsv in enumerate(_tsvs(g, golden_dir)):
            obj.list_mx_info[tsv], list_mxs[tsv] = mod.read_minimizers(tsv)
            weights[tsv] = g["weights"][i]
        gr = mod.build_graph(mod.filter_minimizers(list_mxs), weights)
        obj.args = types.SimpleNamespace(p=f"out_{name}")
        obj.print_graph(gr)
        want = open(os.path.join(golden_dir, "expected", f"print_graph_{name}.mx.dot"), "rb").read()
        assert open(f"out_{name}.mx.dot", "rb").read() == want, name
        scaf.list_mx_info, scaf.graph = obj.list_mx_info, gr
        for target in list_mxs:
            got, ref = scaf.find_mx_min_max(target), original(scaf, target)
            assert got == ref and list(got) == list(ref), (name, target)
    assert not calls
    obj.print_graph(types.SimpleNamespace(vs=[], es=[]))     # not built by the engine: the original must run
    assert len(calls) == 1
    monkeypatch.setattr(dropin, "_LAST_FILTER", None)
