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


def test_python_sketch_matches_oracle_and_goldens(oracle, golden_dir):
    """independent plain-integer restatement of step 1 (no rolling, every window scanned) == oracle/mxo.c, and it
    reproduces the reference's shipped golden tuples under canonical=min"""
    import numpy as np
    import oracle_lib
    from ntjoin_b200 import synth
    # reference golden: tests/expected_outputs/ref.fa.k32.w1000.tsv (hash:pos), legacy min combiner
    names, seq, offs = oracle_lib.read_fasta(os.path.join(golden_dir, "inputs", "ref.fa"))
    got = ref_py.sketch_py(seq, offs, 32, 1000, canonical="min")
    want = open(os.path.join(golden_dir, "expected", "ref.fa.k32.w1000.tsv")).read().split("\t")[1].split()
    assert [f"{h}:{p}" for _, p, h, _, _ in got] == want
    assert ref_py.kmer_hashes_py("ACGT" * 8)[:4] == oracle.kmer_hashes("ACGT" * 8)
    # random records with N runs, IUPAC codes, lower case, a homopolymer and a tandem repeat; sum and min; odd k and w
    rng = np.random.Generator(np.random.PCG64(5))
    seq = synth.random_bases(6000, rng)
    seq[700:760] = ord("N"); seq[1500] = ord("R"); seq[2000:2300] |= 0x20
    seq[3000:3200] = ord("A"); seq[3500:3700] = np.resize(np.frombuffer(b"ACG", dtype=np.uint8), 200)
    offs = np.array([0, 1200, 1230, 1230, 4000, 6000], dtype=np.uint64)
    for k, w, canonical in [(15, 10, "sum"), (21, 33, "sum"), (32, 100, "min"), (9, 4, "sum"), (32, 1000, "sum")]:
        want = oracle.sketch(seq, offs, k, w, canonical=canonical)
        got = ref_py.sketch_py(seq, offs, k, w, canonical=canonical)
        assert [(int(c), int(p), int(h1), int(h0), bool(f)) for c, p, h1, h0, f in
                zip(want["contig"], want["pos"], want["out_hash"], want["min_hash"], want["forward"])] == got, (k, w, canonical)
