"""Experiment: is the state a grow() call leaves behind -- the order of the stack f and of the
adjacency lists -- independent of the sum-rule outcomes of the calls BELOW it?  If it were, the sub-searches of the
reference traversal could be started in parallel from computed states.  It is not: every call is run twice from the
same snapshot, once as it is and once with every deeper check forced to pass, and the two post-states are compared.
Test infrastructure (uses oracle/gm_model.py); python tests/experiments/post_state_experiment.py [seed]"""
import sys, random, copy
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.gm_model import Search, Tree

class S2(Search):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.force_ok_below = None   # depth below which checks are forced true
        self.depth = 0
        self.mism = 0
        self.calls = 0
        self.probe = True
    def check(self, t, n):
        if self.force_ok_below is not None and self.depth > self.force_ok_below:
            return True
        return Search.check(self, t, n)
    def grow(self, t):
        self.depth += 1
        if self.probe and len(t.nodes) < self.n:
            # snapshot, run hypothetical variant, compare with actual afterwards
            snap = (list(self.f), [list(a) for a in self.adj], t.clone(), self.num_grow, list(self.trees), self.L)
            Search.grow(self, t)
            post = (list(self.f), [list(a) for a in self.adj])
            ng, trees, L = self.num_grow, self.trees, self.L
            # variant
            self.f[:] = snap[0]; self.adj = [list(a) for a in snap[1]]
            t2 = snap[2]
            self.trees = []; self.L = snap[5]
            self.probe = False
            self.force_ok_below = self.depth
            Search.grow(self, t2)
            self.force_ok_below = None
            self.probe = True
            post2 = (list(self.f), [list(a) for a in self.adj])
            self.calls += 1
            if post != post2:
                self.mism += 1
            self.f[:] = post[0]; self.adj = [list(a) for a in post[1]]
            self.num_grow, self.trees, self.L = ng, trees, L
        else:
            Search.grow(self, t)
        self.depth -= 1

def rand_instance(rng, n, S, p, dag=True):
    adj = [[] for _ in range(n)]
    for u in range(n):
        for v in range(1, n):
            if u == v: continue
            if dag and not (u < v): continue
            if rng.random() < p or (u == 0 and rng.random() < 0.5):
                adj[u].append(v)
    for a in adj: rng.shuffle(a)
    for v in range(1, n):
        if not any(v in a for a in adj): adj[0].append(v)
    vaf = [[0.5] * S] + [[rng.random() * 0.5 / (1 + 0.3 * v) for _ in range(S)] for v in range(1, n)]
    return adj, vaf

if __name__ == '__main__':
    rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
    tot = mism = 0
    for it in range(60):
        n = rng.randint(4, 8); S = rng.randint(1, 3)
        adj, vaf = rand_instance(rng, n, S, rng.choice([0.3, 0.6, 1.0]), dag=rng.random() < 0.7)
        s = S2(n, S, adj, vaf, 0.02, max_trees=10**9)
        s.get_lineage_trees()
        tot += s.calls; mism += s.mism
        print(it, n, len(s.trees), s.calls, s.mism)
    print('total', tot, 'mismatch', mism)
