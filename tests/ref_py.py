"""Pure-Python restatement of ntJoin's step-2/3 functions (small cases only; TEST INFRASTRUCTURE).

Follows bin/ntjoin_utils.py:167-193 (read_minimizers), :152-165 (filter_minimizers), :83-141
(build_graph) in behaviour; written independently.  Checked against the golden vectors produced by
the reference's own functions (tests/golden/steps23_*.json).  Also provides the recording stand-in
for python-igraph used where igraph is not installed.
"""
import types


class RecordingGraph:
    """Just enough of igraph.Graph for build_graph: names, edges by name or index, es attributes.
    Like the C library it stands in for, it only STORES what add_vertices / add_edges are given (igraph appends to C
    arrays); names and edge ids are resolved when somebody asks.  (The first version did a dict insert and two name
    look-ups per edge in Python: 1.4 s for the 375 k edges of bench.py's file-seam measurement, more than the drop-in
    functions it was standing behind.)"""

    def __init__(self):
        self.vnames, self.eattr = [], {}
        self._raw = []                     # edge end points as given: vertex names or vertex indices
        self._edges = self._eid = None     # resolved views, built on demand

    def add_vertices(self, names):
        self.vnames.extend(names)
        self._edges = self._eid = None

    def add_edges(self, pairs):
        self._raw.extend(pairs)
        self._edges = self._eid = None

    @property
    def edges(self):
        """[(source name, target name)] in insertion order"""
        if self._edges is None:
            vn = self.vnames
            self._edges = [(vn[s] if isinstance(s, int) else s, vn[t] if isinstance(t, int) else t) for s, t in self._raw]
        return self._edges

    def _name(self, v):
        return self.vnames[v] if isinstance(v, int) else v

    def get_eid(self, s, t):
        if self._eid is None:
            self._eid = {}
            for i, (a, b) in enumerate(self.edges):
                self._eid[(a, b)] = self._eid[(b, a)] = i
        return self._eid[(self._name(s), self._name(t))]

    class _ES:
        def __init__(self, g):
            self.g = g

        def __setitem__(self, k, v):
            self.g.eattr[k] = list(v)

        def __getitem__(self, k):
            return self.g.eattr[k]

        def __iter__(self):
            return iter(())

        def __len__(self):
            return len(self.g._raw)

    @property
    def es(self):
        return RecordingGraph._ES(self)

    @property
    def vs(self):
        return [{"name": n} for n in self.vnames]


def read_minimizers(tsv_filename, repeat_bf=False):
    first, dup, per_record = {}, set(), []
    with open(tsv_filename, encoding="utf-8") as fh:
        for raw in fh:
            cols = raw.strip().split("\t")
            if len(cols) < 2:
                continue
            entries = [e.split(":") for e in cols[1].split(" ")]
            per_record.append([e[0] for e in entries])
            for mx, pos, seq in entries:
                if mx in first or (repeat_bf and repeat_bf.contains(seq)):
                    dup.add(mx)
                else:
                    first[mx] = (cols[0], int(pos))
    info = {m: v for m, v in first.items() if m not in dup}
    return info, [[m for m in rec if m not in dup] for rec in per_record]


def filter_minimizers(list_mxs):
    sets = [set(m for rec in recs for m in rec) for recs in list_mxs.values()]
    common = set.intersection(*sets)
    return {asm: [[m for m in rec if m in common] for rec in recs] for asm, recs in list_mxs.items()}


def build_graph(list_mxs, weights, graph=None, black_list=None):
    assert graph is None
    g = RecordingGraph()
    adj, verts = {}, set()
    for asm, recs in list_mxs.items():
        for rec in recs:
            for a, b in zip(rec, rec[1:]):
                if a in adj and b in adj[a]:
                    adj[a][b].append(asm)
                elif b in adj and a in adj[b]:
                    adj[b][a].append(asm)
                else:
                    adj.setdefault(a, {})[b] = [asm]
            for m in rec:
                if black_list is None or m not in black_list:
                    verts.add(m)
    pairs = [(s, t) for s in adj for t in adj[s]]
    g.add_vertices(sorted(verts, key=int))
    g.add_edges(pairs)
    g.es["support"] = [adj[s][t] for s, t in pairs]
    g.es["weight"] = [sum(weights[f] for f in adj[s][t]) for s, t in pairs]
    return g


def as_module():
    """a module object shaped like the reference's ntjoin_utils (for ntjoin_b200.dropin.install)"""
    m = types.ModuleType("ntjoin_utils")
    ig = types.ModuleType("igraph")
    ig.Graph = RecordingGraph
    m.ig = ig
    m.read_minimizers, m.filter_minimizers, m.build_graph = read_minimizers, filter_minimizers, build_graph
    return m


# ---------------------------------------------------------------------------------------------------------
# Step 1 in plain Python integers, written independently of oracle/mxo.c: no rolling, no
# deque -- every valid k-mer is hashed from scratch and every window is scanned.  Small inputs only.
# ---------------------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1
_SEED = {"A": 0x3c8bfbb395c60474, "C": 0x3193c18562a02b4c, "G": 0x20323ed082572324, "T": 0x295549f54be24456}
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def _srol(x, n=1):
    """rotate the 33-bit group (bits 32:0) and the 31-bit group (bits 63:33) left by n, independently"""
    lo, hi = x & ((1 << 33) - 1), x >> 33
    a, b = n % 33, n % 31
    lo = ((lo << a) | (lo >> (33 - a))) & ((1 << 33) - 1) if a else lo
    hi = ((hi << b) | (hi >> (31 - b))) & ((1 << 31) - 1) if b else hi
    return (hi << 33) | lo


def kmer_hashes_py(kmer, canonical="sum"):
    """(fwd, rev, hash0, hash1, forward) of one k-mer of ACGT"""
    k = len(kmer)
    fwd = rev = 0
    for i, ch in enumerate(kmer):
        fwd ^= _srol(_SEED[ch], k - 1 - i)
        rev ^= _srol(_SEED[_COMP[ch]], i)
    h0 = (fwd + rev) & _M64 if canonical == "sum" else min(fwd, rev)
    t = (h0 * (1 ^ ((k * 0x90b45d39fb6da1fa) & _M64))) & _M64
    return fwd, rev, h0, t ^ (t >> 27), fwd <= rev


def sketch_py(seq, offsets, k, w, canonical="sum"):
    """[(record, pos, out_hash, min_hash, forward)] -- windows over VALID k-mers, rightmost minimum, emitted when it moves"""
    text = bytes(seq).decode("latin-1").upper()
    out = []
    for c in range(len(offsets) - 1):
        rec = text[int(offsets[c]):int(offsets[c + 1])]
        valid = []                                   # (pos, hash0, hash1, forward) of every k-mer made of ACGT only
        for p in range(len(rec) - k + 1):
            km = rec[p:p + k]
            if all(ch in _SEED for ch in km):
                _, _, h0, h1, fw = kmer_hashes_py(km, canonical)
                valid.append((p, h0, h1, fw))
        last = -1
        for j in range(w - 1, len(valid)):
            win = valid[j - w + 1:j + 1]
            best = min(x[1] for x in win)
            p, h0, h1, fw = [x for x in win if x[1] == best][-1]      # ties: the rightmost
            if p > last and h0 != _M64:
                out.append((c, p, h1, h0, fw))
                last = p
    return out
