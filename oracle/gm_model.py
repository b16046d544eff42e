"""Pure-Python restatement of LICHeE's lineage-tree search (TEST INFRASTRUCTURE ONLY).

Small-case oracle: follows the reference Java statement by statement, including the
ArrayList mutation order of the frontier stack `f` and of the adjacency lists, because
those decide the enumeration order, which in turn decides (a) the child insertion order
used for the double-precision sums and (b) how Collections.sort (stable) breaks score ties.

Reference (read-only, /root/reference/LICHeE/src/lineage):
  PHYNetwork.grow              PHYNetwork.java:443-541
  PHYNetwork.getLineageTrees   PHYNetwork.java:547-565
  PHYNetwork.evaluateLineageTrees PHYNetwork.java:726-733
  PHYTree.checkConstraint      PHYTree.java:156-174
  PHYTree.computeErrorScore    PHYTree.java:214-233
  PHYTree.addEdge/removeEdge   PHYTree.java:62-96
  PHYTree.isDescendent         PHYTree.java:138-154

Never imported by the product package; only tests/ and scratch experiments use it.
"""
import math
import sys


class Tree:
    """PHYTree: treeNodes (ArrayList), treeEdges (HashMap node -> ArrayList children)."""

    def __init__(self):
        self.nodes = []
        self.edges = {}

    def add_node(self, n):                      # PHYTree.java:56-60
        if n not in self.nodes:
            self.nodes.append(n)

    def add_edge(self, a, b):                   # PHYTree.java:62-70
        l = self.edges.setdefault(a, [])
        if b not in l:
            l.append(b)

    def remove_edge(self, a, b):                # PHYTree.java:72-96
        l = self.edges.get(a)
        if l is not None and b in l:
            l.remove(b)
        connected = any(b in ch for ch in self.edges.values())
        if not connected and b in self.nodes:
            self.nodes.remove(b)

    def clone(self):                            # PHYTree.java:124-133
        c = Tree()
        c.nodes = list(self.nodes)
        c.edges = {k: list(v) for k, v in self.edges.items()}
        return c

    def is_descendent(self, v, w):              # PHYTree.java:138-154
        nb = self.edges.get(v)
        if nb is None:
            return False
        q = list(nb)
        while q:
            n = q.pop(0)
            if n == w:
                return True
            if n in self.edges:
                q.extend(self.edges[n])
        return False


class Search:
    def __init__(self, n_nodes, n_samples, adj, vaf, margin,
                 max_trees=100000, max_grow_calls=100000000, restore_in_place=False):
        """adj: list of child lists in the reference's adjacency order; node 0 is the root.
        vaf[node][sample] == PHYNode.getAAF (root row = VAF_MAX)."""
        self.n = n_nodes
        self.S = n_samples
        self.adj = [list(a) for a in adj]
        self.vaf = vaf
        self.margin = margin
        self.max_trees = max_trees
        self.max_grow = max_grow_calls
        self.trees = []
        self.f = []
        self.L = None
        self.num_grow = 0
        self.restore_in_place = restore_in_place   # experiment switch, NOT the reference
        self.trace = None

    # PHYTree.checkConstraint (PHYTree.java:156-174)
    def check(self, t, n):
        nb = t.edges.get(n)
        if nb is None:
            return True
        for i in range(self.S):
            s = 0.0
            for c in nb:
                s += self.vaf[c][i]
            if s > self.vaf[n][i] + self.margin:
                return False
        return True

    def _remove_g(self, a, b):                  # PHYNetwork.removeEdge :289-299
        if b in self.adj[a]:
            self.adj[a].remove(b)

    def _add_g(self, a, b):                     # PHYNetwork.addEdge :277-286
        if b not in self.adj[a]:
            self.adj[a].append(b)

    def grow(self, t):                          # PHYNetwork.java:443-541
        self.num_grow += 1
        if len(t.nodes) == self.n:
            self.L = t
            self.trees.append(t.clone())
            return
        ff = []
        pos = []
        b = False
        f = self.f
        if self.trace is not None:
            self.trace_enter(t)
        while (not b) and f:
            e = f.pop()
            u, v = e
            t.add_node(v)
            t.add_edge(u, v)
            ok = self.check(t, u)
            if self.trace is not None:
                self.trace_pop(t, e, ok)
            if ok:
                added = []
                for w in self.adj[v]:
                    if w not in t.nodes:
                        f.append((v, w))
                        added.append((v, w))
                removed = [x for x in f if (x[0] in t.nodes and x[1] == v)]
                f[:] = [x for x in f if x not in removed]
                if self.num_grow >= self.max_grow:
                    return
                self.grow(t)
                if len(self.trees) == self.max_trees:
                    return
                f[:] = [x for x in f if x not in added]
                f.extend(removed)
            t.remove_edge(u, v)
            if self.restore_in_place:
                pos.append(self.adj[u].index(v))
            self._remove_g(u, v)
            ff.append(e)
            # bridge test (PHYNetwork.java:515-530)
            b = True
            for w in range(self.n):
                if v in self.adj[w]:
                    if self.L is None or not self.L.is_descendent(v, w):
                        b = False
                        break
        for i in range(len(ff) - 1, -1, -1):
            e = ff[i]
            f.append(e)
            if self.restore_in_place:
                self.adj[e[0]].insert(pos[i], e[1])
            else:
                self._add_g(e[0], e[1])

    def trace_enter(self, t):
        pass

    def trace_pop(self, t, e, ok):
        pass

    def get_lineage_trees(self):                # PHYNetwork.java:547-565
        self.trees = []
        t = Tree()
        t.add_node(0)
        self.f = []
        if not self.adj[0]:
            return self.trees
        for n in self.adj[0]:
            self.f.append((0, n))
        old = sys.getrecursionlimit()
        sys.setrecursionlimit(max(old, 10 * self.n + 1000))
        self.grow(t)
        return self.trees

    # PHYTree.computeErrorScore (PHYTree.java:214-233); PHYNode.compareTo sorts ids descending
    def score(self, t):
        err = 0.0
        for n in sorted(t.edges.keys(), reverse=True):
            nb = t.edges[n]
            for i in range(self.S):
                s = 0.0
                for c in nb:
                    s += self.vaf[c][i]
                if s > self.vaf[n][i]:
                    d = s - self.vaf[n][i]
                    err += d * d                 # Math.pow(x, 2) == x*x (fdlibm special case)
        return math.sqrt(err)

    def evaluate(self):                          # PHYNetwork.java:726-733, stable sort
        scored = [(self.score(t), i, t) for i, t in enumerate(self.trees)]
        scored.sort(key=lambda x: x[0])          # Python sort is stable like Collections.sort
        return scored


def tree_key(t):
    """Canonical content of a tree: tuple of (parent, children-in-insertion-order)."""
    return tuple(sorted((p, tuple(ch)) for p, ch in t.edges.items() if ch))


def parent_vec(t, n):
    par = [-1] * n
    for p, ch in t.edges.items():
        for c in ch:
            par[c] = p
    return tuple(par)
