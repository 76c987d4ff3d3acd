"""CPU: the pure-Python restatement of steps 2-3 reproduces the reference-generated golden vectors."""
import glob
import json
import os
import subprocess

import oracle_lib
import ref_py


def test_ref_py_vs_golden(golden_dir, tmp_path):
    oracle_lib.build()
    for path in sorted(glob.glob(os.path.join(golden_dir, "steps23_*.json"))):
        g = json.load(open(path))
        list_mxs, infos, weights = {}, [], {}
        for i, f in enumerate(g["files"]):
            tsv = str(tmp_path / f"{i}.{f}.tsv")
            subprocess.check_call([oracle_lib.CLI, "--seq", "--long", "--pos", "-k", str(g["k"]), "-w", str(g["w"]),
                                   os.path.join(golden_dir, "inputs", f), "-o", tsv])
            info, mxs = ref_py.read_minimizers(tsv)
            assert {k: list(v) for k, v in info.items()} == g["read_minimizers"][i]["mx_info"]
            assert list(info) == g["read_minimizers"][i]["mx_order"]      # same insertion order
            assert mxs == g["read_minimizers"][i]["mxs"]
            list_mxs[tsv], weights[tsv] = mxs, g["weights"][i]
        filt = ref_py.filter_minimizers(list_mxs)
        assert [filt[t] for t in list_mxs] == g["filter_minimizers"]
        gr = ref_py.build_graph(filt, weights)
        keys = list(list_mxs)
        assert [list(e) for e in gr.edges] == g["edges"]
        assert [[keys.index(f) for f in s] for s in gr.eattr["support"]] == g["support"]
        assert gr.eattr["weight"] == g["weight"] and gr.vnames == g["vertices"]
