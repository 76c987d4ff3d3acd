"""CPU, build container only (needs /root/reference): the oracle's steps 2-3 against the reference's OWN functions
read_minimizers / filter_minimizers / build_graph (bin/ntjoin_utils.py, imported unmodified) on RANDOM minimizer lists:
hashes drawn from a small universe, so that duplicates inside an assembly, hashes missing from some assemblies, records
that lose all their minimizers, edges seen in several assemblies in either orientation and sources with several edges all
occur.  The fifteen committed golden cases come from real sequence, where most of that is rare; this test is what pins the
edge ORDER (formatted_edges, bin/ntjoin_utils.py:115), the support lists and the weights for 1 to 5 assemblies.
The engine is compared with the oracle on the GPU (tests/test_gpu_filter.py); this closes the other half of the chain."""
import importlib.util
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "bin")), reason="needs the reference checkout")


@pytest.fixture(scope="module")
def ref_utils():
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    saved = sys.modules.get("igraph")
    utils = mg.load_reference_utils()
    import ntjoin as ntjoin_mod                    # the reference's bin/ntjoin.py (print_graph)
    utils._mg, utils._ntjoin_mod = mg, ntjoin_mod
    yield utils
    if saved is not None:
        sys.modules["igraph"] = saved
    else:
        sys.modules.pop("igraph", None)


def _random_case(rng, n_asm, universe, n_rec, max_len):
    """per assembly: list of records, each a list of hashes (ints)"""
    pool = rng.integers(1, 2**63, size=universe, dtype=np.int64).astype(np.uint64) * np.uint64(2) + np.uint64(1)
    pool = np.unique(pool)
    base = [pool[rng.permutation(len(pool))[:rng.integers(2, max_len + 1)]] for _ in range(n_rec)]      # "true" records
    asms = []
    for _a in range(n_asm):
        recs = []
        for b in base:
            r = b.copy()
            if rng.random() < 0.5 and len(r) > 3:                # a misassembly: pieces in another order / orientation, so that
                c1, c2 = sorted(rng.integers(1, len(r), size=2))  # the assemblies disagree about who is next to whom
                pieces = [p for p in (r[:c1], r[c1:c2], r[c2:]) if len(p)]
                pieces = [pieces[i] for i in rng.permutation(len(pieces))]
                r = np.concatenate([p[::-1] if rng.random() < 0.5 else p for p in pieces])
            if rng.random() < 0.5:
                r = r[::-1]                                      # the same record in the other orientation
            if rng.random() < 0.4:
                cut = int(rng.integers(1, len(r)))               # broken into two contigs
                recs += [r[:cut], r[cut:]]
            else:
                recs.append(r)
        for _d in range(int(rng.integers(0, 4))):                # a few extra records: duplicates and foreign hashes
            recs.append(pool[rng.integers(0, len(pool), size=int(rng.integers(1, 6)))])
        order = rng.permutation(len(recs))
        asms.append([recs[i] for i in order])
    return asms


def _write_tsv(path, recs):
    with open(path, "w") as fh:
        for c, r in enumerate(recs):
            fh.write(f"ctg{c}\t" + " ".join(f"{int(h)}:{7 * i + c}:ACGT" for i, h in enumerate(r)) + "\n")


@pytest.mark.parametrize("n_asm,universe,n_rec,max_len,seed", [(1, 40, 3, 12, 0), (2, 60, 4, 15, 1), (2, 25, 5, 10, 2), (3, 80, 5, 20, 3),
                                                              (3, 30, 6, 8, 4), (4, 100, 6, 18, 5), (5, 70, 5, 14, 6), (2, 400, 12, 60, 7)])
def test_random_lists_oracle_equals_reference(oracle, ref_utils, tmp_path, monkeypatch, n_asm, universe, n_rec, max_len, seed):
    import ref_py
    from ntjoin_b200 import dropin
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "harness"))
    from fake_engine import FakeEngine
    monkeypatch.setattr(dropin, "_ENGINE", FakeEngine())
    monkeypatch.setattr(dropin, "_LAST_FILTER", None)
    dropin_mod = dropin.install(ref_py.as_module())
    rng = np.random.default_rng(1000 + seed)
    weights_all = [2.0, 1.0, 1.5, 0.1, 3.0]
    seen = {"edges": 0, "multi_support": 0, "not_unique": 0, "dropped_by_intersection": 0, "sources_with_several_edges": 0}
    for rep in range(20):
        asms = _random_case(rng, n_asm, universe, n_rec, max_len)
        tsvs, list_mxs, wdict = [], {}, {}
        with ref_utils.HiddenPrints():
            for a, recs in enumerate(asms):
                tsv = str(tmp_path / f"{seed}.{rep}.{a}.tsv")
                _write_tsv(tsv, recs)
                tsvs.append(tsv)
                _info, list_mxs[tsv] = ref_utils.read_minimizers(tsv)
                wdict[tsv] = weights_all[a]
            filtered = ref_utils.filter_minimizers(list_mxs)
            graph = ref_utils.build_graph(filtered, wdict)
        hashes = [np.concatenate(recs) for recs in asms]
        contigs = [np.concatenate([np.full(len(r), c, dtype=np.uint32) for c, r in enumerate(recs)]) for recs in asms]
        got = oracle.filter_and_edges(hashes, contigs, weights_all[:n_asm])
        for a, tsv in enumerate(tsvs):
            bounds = np.cumsum([0] + [len(r) for r in asms[a]])
            uniq_lists = [[str(int(h)) for h in hashes[a][s:e][got["uniq"][a][s:e]]] for s, e in zip(bounds[:-1], bounds[1:])]
            keep_lists = [[str(int(h)) for h in hashes[a][s:e][got["keep"][a][s:e]]] for s, e in zip(bounds[:-1], bounds[1:])]
            assert uniq_lists == list_mxs[tsv], (seed, rep, a)
            assert keep_lists == filtered[tsv], (seed, rep, a)
        assert sorted(graph.vnames, key=int) == [str(int(v)) for v in got["vertices"]], (seed, rep)
        e = got["edges"]
        assert [[s, t] for s, t in graph.edges] == [[str(int(u)), str(int(v))] for u, v in zip(e["u"], e["v"])], (seed, rep)
        key = {t: i for i, t in enumerate(tsvs)}
        assert [sum(1 << key[f] for f in sup) for sup in graph.eattr["support"]] == e["support_mask"].tolist(), (seed, rep)
        assert [[key[f] for f in sup] for sup in graph.eattr["support"]] == \
            [[b for b in range(n_asm) if m >> b & 1] for m in e["support_mask"].tolist()], (seed, rep)     # support lists in assembly order
        assert graph.eattr["weight"] == e["weight"].tolist(), (seed, rep)
        # the product's .mx.dot writer (host only) from arrays, against the reference's own Ntjoin.print_graph
        if rep % 4 == 0:
            from ntjoin_b200.dot import write_mx_dot
            list_mx_info = {}
            with ref_utils.HiddenPrints():
                for tsv in tsvs:
                    list_mx_info[tsv], _m = ref_utils.read_minimizers(tsv)
            vnames = [str(int(v)) for v in got["vertices"]]
            base = {t: os.path.basename(t) for t in tsvs}
            want_path = str(tmp_path / f"want.{seed}.{rep}.mx.dot")
            ref_utils._mg.reference_print_graph(ref_utils._ntjoin_mod, ref_utils, vnames, graph.edges,
                                                [[base[f] for f in sup] for sup in graph.eattr["support"]], graph.eattr["weight"],
                                                {base[t]: list_mx_info[t] for t in tsvs}, want_path)
            index = {v: i for i, v in enumerate(vnames)}
            names, v_ctg, v_pos = [], [], []
            for t in tsvs:
                info = list_mx_info[t]
                nm = sorted({c for c, _p in info.values()})
                at = {c: i for i, c in enumerate(nm)}
                names.append(nm)
                v_ctg.append([at[info[v][0]] for v in vnames])
                v_pos.append([info[v][1] for v in vnames])
            pairs = [sorted((index[str(int(u))], index[str(int(v))])) for u, v in zip(e["u"], e["v"])]   # igraph: source = lower id
            got_path = tmp_path / f"got.{seed}.{rep}.mx.dot"
            write_mx_dot(got_path, got["vertices"], [base[t] for t in tsvs], names, v_ctg, v_pos,
                         [p[0] for p in pairs], [p[1] for p in pairs], e["support_mask"], weights_all[:n_asm])
            want_bytes = open(want_path, "rb").read()
            assert got_path.read_bytes() == want_bytes, (seed, rep)
            assert want_bytes.count(b" --") == len(e) and want_bytes.count(b"[label=") == len(vnames), (seed, rep)
        # the product's drop-in functions (ntjoin_b200/dropin.py, engine calls served by the oracle-backed stand-in)
        # against the reference's: same dict (content and insertion order), same lists, same graph
        if rep % 2 == 1:
            d_info, d_mxs = {}, {}
            for tsv in tsvs:
                d_info[tsv], d_mxs[tsv] = dropin_mod.read_minimizers(tsv)
            with ref_utils.HiddenPrints():
                for tsv in tsvs:
                    r_info, _m = ref_utils.read_minimizers(tsv)
                    assert d_info[tsv] == r_info and list(d_info[tsv]) == list(r_info), (seed, rep)
            assert {t: list(v) for t, v in d_mxs.items()} == list_mxs, (seed, rep)
            d_filt = dropin_mod.filter_minimizers(d_mxs)
            assert {t: list(v) for t, v in d_filt.items()} == filtered, (seed, rep)
            d_graph = dropin_mod.build_graph(d_filt, wdict)
            assert sorted(d_graph.vnames, key=int) == sorted(graph.vnames, key=int), (seed, rep)
            assert d_graph.edges == graph.edges, (seed, rep)
            assert d_graph.eattr["support"] == graph.eattr["support"] and d_graph.eattr["weight"] == graph.eattr["weight"], (seed, rep)
        seen["edges"] += len(e)
        seen["multi_support"] += int(sum(bin(m).count("1") > 1 for m in e["support_mask"].tolist()))
        seen["not_unique"] += int(sum((~u).sum() for u in got["uniq"]))
        seen["dropped_by_intersection"] += int(sum((u & ~k).sum() for u, k in zip(got["uniq"], got["keep"])))
        srcs = [s for s, _t in graph.edges]
        seen["sources_with_several_edges"] += len(srcs) - len(set(srcs))
    monkeypatch.setattr(dropin, "_LAST_FILTER", None)
    # the random cases must really contain what this test is about
    assert seen["edges"] > 20 and seen["not_unique"] > 10, seen
    if n_asm > 1:
        assert seen["multi_support"] > 10 and seen["dropped_by_intersection"] > 10 and seen["sources_with_several_edges"] > 0, seen
        assert seen["edges"] - seen["multi_support"] > 0, seen          # edges that only some assemblies support
