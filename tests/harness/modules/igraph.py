"""python-igraph for the reference harness: the undirected-Graph subset ntJoin calls
(bin/ntjoin_utils.py:37-141, bin/ntjoin.py:43-167, bin/ntjoin_assemble.py:78-81,693, bin/ntjoin_overlap.py:29-43,139-158).

Conventions kept from igraph: vertex and edge ids are dense and renumbered after deletions; an undirected edge
reports the lower vertex id as its source; subgraph() keeps the relative order of vertices and edges; components are
listed by their lowest vertex id; vs.find(str) looks a vertex up by name and raises ValueError when it is missing.
"""
from collections import deque


class Vertex:
    __slots__ = ("graph", "index")

    def __init__(self, graph, index):
        self.graph, self.index = graph, index

    def __getitem__(self, key):
        return self.graph._vattr[key][self.index]

    def degree(self):
        return len(self.graph._adj()[self.index])


class Edge:
    __slots__ = ("graph", "index")

    def __init__(self, graph, index):
        self.graph, self.index = graph, index

    source = property(lambda self: min(self.graph._edges[self.index]))
    target = property(lambda self: max(self.graph._edges[self.index]))
    tuple = property(lambda self: (self.source, self.target))

    def __getitem__(self, key):
        return self.graph._eattr[key][self.index]


class VertexSeq:
    def __init__(self, graph):
        self.graph = graph

    def __call__(self):
        return self

    def __len__(self):
        return len(self.graph._vattr["name"])

    def __iter__(self):
        return (Vertex(self.graph, i) for i in range(len(self)))

    def __getitem__(self, key):
        if isinstance(key, str):
            return list(self.graph._vattr[key])
        if key < 0 or key >= len(self):
            raise IndexError("vertex index out of range")
        return Vertex(self.graph, key)

    def find(self, key):
        if isinstance(key, int):
            return self[key]
        try:
            return Vertex(self.graph, self.graph._index()[key])
        except KeyError:
            raise ValueError(f"no such vertex: {key!r}") from None


class EdgeSeq:
    def __init__(self, graph, ids=None):
        self.graph, self.ids = graph, ids

    def __call__(self):
        return self

    def _all(self):
        return range(len(self.graph._edges)) if self.ids is None else self.ids

    def __len__(self):
        return len(self._all())

    def __iter__(self):
        return (Edge(self.graph, i) for i in self._all())

    def __getitem__(self, key):
        if isinstance(key, str):
            return [self.graph._eattr[key][i] for i in self._all()]
        if isinstance(key, (list, tuple, set, range)):
            return EdgeSeq(self.graph, [self._all()[i] if self.ids is not None else i for i in key])
        return Edge(self.graph, self._all()[key])

    def __setitem__(self, key, values):
        values = list(values)
        ids = list(self._all())
        if len(values) != len(ids):
            raise ValueError("attribute list length must match the number of edges")
        col = self.graph._eattr.setdefault(key, [None] * len(self.graph._edges))
        for i, v in zip(ids, values):
            col[i] = v


class Graph:
    def __init__(self):
        self._vattr = {"name": []}
        self._edges = []             # (u, v) vertex ids as given
        self._eattr = {}
        self._gattr = {}             # graph attributes: g["key"] = value, kept by copy() and subgraph()
        self._cache = {}

    def __setitem__(self, key, value):
        self._gattr[key] = value

    def __getitem__(self, key):
        return self._gattr[key]

    def attributes(self):
        return list(self._gattr)

    # ---- internals
    def _touch(self):
        self._cache = {}

    def _index(self):
        if "index" not in self._cache:
            self._cache["index"] = {n: i for i, n in enumerate(self._vattr["name"])}
        return self._cache["index"]

    def _adj(self):
        if "adj" not in self._cache:
            adj = [[] for _ in self._vattr["name"]]
            for eid, (u, v) in enumerate(self._edges):
                adj[u].append((v, eid))
                if v != u:
                    adj[v].append((u, eid))
            self._cache["adj"] = adj
        return self._cache["adj"]

    def _vid(self, v):
        if isinstance(v, Vertex):
            return v.index
        if isinstance(v, str):
            try:
                return self._index()[v]
            except KeyError:
                raise ValueError(f"no such vertex: {v!r}") from None
        return int(v)

    # ---- construction
    def add_vertices(self, names):
        if isinstance(names, int):
            names = [None] * names
        self._vattr["name"].extend(names)
        self._touch()

    def add_edges(self, pairs):
        new = [(self._vid(s), self._vid(t)) for s, t in pairs]
        self._edges.extend(new)
        for col in self._eattr.values():
            col.extend([None] * len(new))
        self._touch()

    def copy(self):
        g = Graph()
        g._vattr = {k: list(v) for k, v in self._vattr.items()}
        g._edges = list(self._edges)
        g._eattr = {k: list(v) for k, v in self._eattr.items()}
        g._gattr = dict(self._gattr)
        return g

    def delete_edges(self, ids):
        if isinstance(ids, (int, Edge)):
            ids = [ids]
        drop = {e.index if isinstance(e, Edge) else int(e) for e in ids}
        keep = [i for i in range(len(self._edges)) if i not in drop]
        self._edges = [self._edges[i] for i in keep]
        self._eattr = {k: [v[i] for i in keep] for k, v in self._eattr.items()}
        self._touch()

    # ---- queries
    vs = property(lambda self: VertexSeq(self))
    es = property(lambda self: EdgeSeq(self))

    def vcount(self):
        return len(self._vattr["name"])

    def ecount(self):
        return len(self._edges)

    def incident(self, v):
        return [eid for _, eid in self._adj()[self._vid(v)]]

    def neighbors(self, v):
        return [u for u, _ in self._adj()[self._vid(v)]]

    def get_eid(self, a, b):
        a, b = self._vid(a), self._vid(b)
        for u, eid in self._adj()[a]:
            if u == b:
                return eid
        raise ValueError(f"no such edge: {a} -- {b}")

    def components(self):
        seen, out = [False] * self.vcount(), []
        adj = self._adj()
        for s in range(self.vcount()):
            if seen[s]:
                continue
            seen[s] = True
            comp, queue = [], deque([s])
            while queue:
                u = queue.popleft()
                comp.append(u)
                for v, _ in adj[u]:
                    if not seen[v]:
                        seen[v] = True
                        queue.append(v)
            out.append(sorted(comp))
        return out

    def subgraph(self, vertices):
        ids = sorted({self._vid(v) for v in vertices})
        new = {old: i for i, old in enumerate(ids)}
        g = Graph()
        g._vattr = {k: [v[i] for i in ids] for k, v in self._vattr.items()}
        keep = [i for i, (u, v) in enumerate(self._edges) if u in new and v in new]
        g._edges = [(new[self._edges[i][0]], new[self._edges[i][1]]) for i in keep]
        g._eattr = {k: [v[i] for i in keep] for k, v in self._eattr.items()}
        g._gattr = dict(self._gattr)
        return g

    def get_shortest_paths(self, v, to=None, output="vpath", **_kw):
        src = self._vid(v)
        adj = self._adj()
        prev = {src: None}
        queue = deque([src])
        while queue:
            u = queue.popleft()
            for x, _ in adj[u]:
                if x not in prev:
                    prev[x] = u
                    queue.append(x)
        if to is None:
            targets = range(self.vcount())
        elif isinstance(to, (list, tuple)):
            targets = [self._vid(t) for t in to]
        else:
            targets = [self._vid(to)]
        paths = []
        for t in targets:
            path = []
            if t in prev:
                while t is not None:
                    path.append(t)
                    t = prev[t]
                path.reverse()
            paths.append(path)
        return paths
