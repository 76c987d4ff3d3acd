"""Oracle-backed stand-in for ntjoin_b200.Engine (TEST INFRASTRUCTURE, CPU only).

Lets tests/test_reference_suite.py run the reference's own test-suite with the PRODUCT's drop-in layer switched on
(ntjoin_b200/dropin.py, ntjoin_b200/dot.py, dropin/sitecustomize.py: read_minimizers / filter_minimizers / build_graph /
print_graph / find_mx_min_max) in a container without a GPU: only the two engine calls the drop-in layer makes
(load_tsv, filter_and_edges) are served here, by the CPU oracle.  It is injected from outside by
tests/harness/dropin_site/sitecustomize.py; the product has no hook for it and never imports it.
"""
import numpy as np

import oracle_lib


class FakeSketch:
    def __init__(self, names, out_hash, pos, contig):
        self.names = names
        self.out_hash = np.asarray(out_hash, dtype=np.uint64)
        self.pos = np.asarray(pos, dtype=np.uint32)
        self.contig = np.asarray(contig, dtype=np.uint32)
        self.n = len(self.out_hash)

    def close(self):
        pass


class FakeResult:
    def __init__(self, d):
        self.uniq, self.keep, self.vertices = d["uniq"], d["keep"], d["vertices"]
        e = d["edges"]
        self.edge_u, self.edge_v, self.support, self.weight = e["u"].copy(), e["v"].copy(), e["support_mask"].copy(), e["weight"].copy()

    def close(self):
        pass


class FakeEngine:
    def __init__(self):
        self._oracle = oracle_lib.Oracle()

    def load_tsv(self, path):
        """same record / entry rules as mxe_sketch_load_tsv (every line is a record; hash[:pos[:...]] entries)"""
        names, hashes, pos, contig = [], [], [], []
        with open(path, encoding="utf-8") as fh:
            for line in fh:
                line = line.rstrip("\n")
                ident, _, rest = line.partition("\t")
                c = len(names)
                names.append(ident.rstrip("\r "))
                for entry in rest.split():
                    f = entry.split(":")
                    hashes.append(int(f[0]))
                    pos.append(int(f[1]) if len(f) > 1 else 0)
                    contig.append(c)
        return FakeSketch(names, hashes, pos, contig)

    def filter_and_edges(self, sketches, weights):
        return FakeResult(self._oracle.filter_and_edges([s.out_hash for s in sketches], [s.contig for s in sketches], list(weights)))
