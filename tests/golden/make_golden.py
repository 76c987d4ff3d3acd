#!/usr/bin/env python3
"""Generates the committed golden fixtures under tests/golden/ from the reference checkout.

Run in the build container only (needs /root/reference; the GPU box never runs this):
    python tests/golden/make_golden.py

What it records
  inputs/*.fa            small FASTA fixtures of the reference's own test-suite (tests/*.fa), verbatim
  expected/*.tsv,*.dot   the reference's shipped golden outputs (tests/expected_outputs/)
  expected/print_graph_*.mx.dot   the reference's OWN Ntjoin.print_graph (bin/ntjoin.py:25-67) run on the graph of
                         each steps23 case (vertices in ascending-hash order, edge endpoints as igraph reports them:
                         lower vertex id first), with the assembly keys "<i>.<fasta>.k<k>.w<w>.tsv"
  steps23_*.json         outputs of the reference's OWN Python functions read_minimizers,
                         filter_minimizers and build_graph (bin/ntjoin_utils.py, imported unmodified
                         from /root/reference/bin with a recording stand-in for python-igraph), run on
                         TSVs produced by the oracle CLI with `--seq --pos` exactly as ntJoin:205 does.
  sketch_digests.json    oracle digests on tests/ref.longer.fa cross-checked against SURVEY.md 8(c)
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402

SMALL = ["ref.fa", "ref.multiple.fa", "scaf.f-f.fa", "scaf.f-f.copy.fa", "scaf.f-r.fa", "scaf.r-f.fa", "scaf.r-r.fa",
         "scaf.f-f.termN.fa", "scaf.f-f.termN.unassigned.fa", "scaf.multiple.fa", "scaf.misassembled.f-f.r-r.fa",
         "scaf.misassembled.f-r.r-f.fa", "scaf.f-f.overlapping.fa", "scaf.more_seqs.fa"]

# (name, [reference fastas...], target fasta, k, w, weights refs..., target)
CASES = [
    ("config1_ff_w500", ["ref.fa"], "scaf.f-f.fa", 32, 500, [2.0, 1.0]),
    ("ff_w1000", ["ref.fa"], "scaf.f-f.fa", 32, 1000, [2.0, 1.0]),
    ("three_way_w1000", ["ref.fa", "scaf.f-f.copy.fa"], "scaf.f-f.fa", 32, 1000, [2.0, 2.0, 1.0]),
    ("multiple_w500", ["ref.multiple.fa"], "scaf.multiple.fa", 32, 500, [2.0, 1.0]),
    ("misassembled_ffrr_w500", ["ref.multiple.fa"], "scaf.misassembled.f-f.r-r.fa", 32, 500, [2.0, 1.0]),
    ("misassembled_frrf_w500", ["ref.multiple.fa"], "scaf.misassembled.f-r.r-f.fa", 32, 500, [2.0, 1.0]),
    ("termN_w1000", ["ref.fa"], "scaf.f-f.termN.unassigned.fa", 32, 1000, [2.0, 1.0]),
    ("selfdup_w250", ["scaf.more_seqs.fa"], "scaf.more_seqs.fa", 32, 250, [2.0, 1.0]),
    ("overlap_k15_w10", ["ref.fa"], "scaf.f-f.overlapping.fa", 15, 10, [1.0, 1.0]),
    ("rr_w500", ["ref.fa"], "scaf.r-r.fa", 32, 500, [2.0, 1.0]),
    ("fr_w1000", ["ref.fa"], "scaf.f-r.fa", 32, 1000, [2.0, 1.0]),
    ("four_way_w500", ["ref.fa", "scaf.f-f.copy.fa", "scaf.r-f.fa"], "scaf.f-f.fa", 32, 500, [2.0, 2.0, 1.5, 1.0]),
    ("multiple_k24_w250", ["ref.multiple.fa"], "scaf.multiple.fa", 24, 250, [2.0, 1.0]),
    ("termN_k40_w500", ["ref.fa"], "scaf.f-f.termN.fa", 40, 500, [3.0, 1.0]),
    ("more_seqs_vs_misassembled_k20_w50", ["scaf.more_seqs.fa"], "scaf.misassembled.f-f.r-r.fa", 20, 50, [1.0, 1.0]),
]


class _Seq:
    def __init__(self, rows, attrs):
        self.rows, self.attrs = rows, attrs

    def __setitem__(self, key, values):
        self.attrs[key] = list(values)

    def __iter__(self):
        return iter(self.rows)


class FakeGraph:
    """Records what build_graph hands to python-igraph (bin/ntjoin_utils.py:117-141)."""

    def __init__(self):
        self.vnames, self.edges, self.eattr = [], [], {}
        self._eid = {}

    def add_vertices(self, names):
        self.vnames.extend(names)

    def add_edges(self, pairs):
        for s, t in pairs:
            self._eid[(s, t)] = len(self.edges)
            self._eid[(t, s)] = len(self.edges)
            self.edges.append((s, t))

    def get_eid(self, s, t):
        return self._eid[(s, t)]

    def es(self):
        return _Seq(list(range(len(self.edges))), self.eattr)

    def vs(self):
        return _Seq([{"name": n} for n in self.vnames], {})


class DotGraph:
    """What Ntjoin.print_graph reads from python-igraph: vs() / vs[i]['name'], es() with source / target (an
    undirected igraph edge reports the lower vertex id as its source) and the weight / support attributes."""

    class _Edge(dict):
        pass

    class _VS(list):
        def __call__(self):
            return self

    def __init__(self, names, edges, support, weight):
        index = {n: i for i, n in enumerate(names)}
        self.vs = DotGraph._VS({"name": n} for n in names)
        self._es = []
        for (s, t), sup, w in zip(edges, support, weight):
            e = DotGraph._Edge(weight=w, support=sup)
            e.source, e.target = sorted((index[s], index[t]))
            self._es.append(e)

    def es(self):
        return self._es


def reference_print_graph(ntjoin_mod, utils, names, edges, support, weight, list_mx_info, out_path):
    obj = ntjoin_mod.Ntjoin.__new__(ntjoin_mod.Ntjoin)
    obj.list_mx_info = list_mx_info
    obj.args = types.SimpleNamespace(p=out_path[:-len(".mx.dot")])
    with utils.HiddenPrints():
        obj.print_graph(DotGraph(names, edges, support, weight))


def load_reference_utils():
    ig = types.ModuleType("igraph")
    ig.Graph = FakeGraph
    sys.modules["igraph"] = ig
    sys.path.insert(0, os.path.join(REF, "bin"))
    import ntjoin_utils
    return ntjoin_utils


def main():
    oracle_lib.build()
    os.makedirs(os.path.join(HERE, "inputs"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "expected"), exist_ok=True)
    for f in SMALL:
        shutil.copyfile(os.path.join(REF, "tests", f), os.path.join(HERE, "inputs", f))
    for f in ["ref.fa.k32.w1000.tsv", "scaf.f-f.fa.k32.w1000.tsv", "f-f_test.mx.dot"]:
        shutil.copyfile(os.path.join(REF, "tests", "expected_outputs", f), os.path.join(HERE, "expected", f))

    utils = load_reference_utils()
    import ntjoin as ntjoin_mod            # the reference's bin/ntjoin.py (igraph is the recording stand-in)
    for name, refs, target, k, w, weights in CASES:
        with tempfile.TemporaryDirectory() as tmp:
            files = refs + [target]          # assembly order: references, then target (ntjoin_assemble.py:804-807)
            tsvs = []
            for i, f in enumerate(files):
                tsv = os.path.join(tmp, f"{i}.{f}.k{k}.w{w}.tsv")
                subprocess.check_call([oracle_lib.CLI, "--seq", "--long", "--pos", "-k", str(k), "-w", str(w), "-t", "1",
                                       os.path.join(REF, "tests", f), "-o", tsv])
                tsvs.append(tsv)
            list_mx_info, list_mxs, wdict = {}, {}, {}
            with utils.HiddenPrints():
                for tsv, wt in zip(tsvs, weights):
                    info, mxs = utils.read_minimizers(tsv)
                    list_mx_info[tsv], list_mxs[tsv], wdict[tsv] = info, mxs, wt
                filtered = utils.filter_minimizers(list_mxs)
                graph = utils.build_graph(filtered, wdict)
            key = {t: i for i, t in enumerate(tsvs)}
            out = {
                "files": files, "k": k, "w": w, "weights": weights,
                "read_minimizers": [{"mx_info": {mx: [c, p] for mx, (c, p) in list_mx_info[t].items()},
                                     "mx_order": list(list_mx_info[t]),       # dict insertion order
                                     "mxs": list_mxs[t]} for t in tsvs],
                "filter_minimizers": [filtered[t] for t in tsvs],
                "vertices": sorted(graph.vnames, key=int),
                "edges": [[s, t] for s, t in graph.edges],
                "support": [[key[f] for f in sup] for sup in graph.eattr["support"]],
                "weight": graph.eattr["weight"],
            }
            with open(os.path.join(HERE, f"steps23_{name}.json"), "w") as fh:
                json.dump(out, fh, indent=0, sort_keys=True)
            base = {t: os.path.basename(t) for t in tsvs}
            reference_print_graph(ntjoin_mod, utils, out["vertices"], graph.edges,
                                  [[base[f] for f in sup] for sup in graph.eattr["support"]], graph.eattr["weight"],
                                  {base[t]: list_mx_info[t] for t in tsvs},
                                  os.path.join(HERE, "expected", f"print_graph_{name}.mx.dot"))
            print(name, "vertices", len(out["vertices"]), "edges", len(out["edges"]))

    # sketch digests on the large reference fixture (not copied: 13.8 MB)
    orc = oracle_lib.Oracle()
    names, seq, offs = oracle_lib.read_fasta(os.path.join(REF, "tests", "ref.longer.fa"))
    dig = {}
    for w in (1000, 500, 250):
        m = orc.sketch(seq, offs, 32, w)
        dig[f"ref.longer.fa.k32.w{w}"] = {
            "n": int(len(m)), "xor_out_hash": hex(int(np.bitwise_xor.reduce(m["out_hash"]))),
            "sum_pos": int(m["pos"].sum()), "first": [int(m["out_hash"][0]), int(m["pos"][0])],
            "last": [int(m["out_hash"][-1]), int(m["pos"][-1])]}
    with open(os.path.join(HERE, "sketch_digests.json"), "w") as fh:
        json.dump(dig, fh, indent=1, sort_keys=True)
    print(dig)


if __name__ == "__main__":
    main()
