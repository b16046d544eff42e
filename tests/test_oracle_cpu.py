"""CPU tests of the oracle (no GPU): the C restatement against the independent Python restatement,
against the committed golden fixtures, and against the two top trees the reference publishes."""
import json
import os
import random

import numpy as np
import pytest

import oracle
from oracle.gm_model import Search, parent_vec

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rand_dag(rng, n, S, p_edge, vaf_scale):
    order = list(range(1, n)); rng.shuffle(order); order = [0] + order
    rank = {v: i for i, v in enumerate(order)}
    adj = [[] for _ in range(n)]
    for i in range(n):
        for j in range(i + 1, n):
            if rng.random() < p_edge:
                adj[order[i]].append(order[j])
    for v in range(1, n):
        if not any(v in a for a in adj):
            adj[order[rng.randrange(rank[v])]].append(v)
    for a in adj:
        rng.shuffle(a)
    vaf = np.array([[0.5] * S] + [[rng.random() * vaf_scale for _ in range(S)] for _ in range(n - 1)])
    return adj, vaf


@pytest.fixture(scope="module")
def networks():
    return json.load(open(os.path.join(GOLD, "networks_bundled.json")))


def test_c_port_equals_python_port_including_caps():
    """Two independent restatements of PHYNetwork.grow (C and Python) agree on enumeration order,
    child insertion order, scores, stable ranking, numGrowCalls and both early-return caps."""
    rng = random.Random(3)
    for _ in range(120):
        n = rng.randint(2, 8); S = rng.randint(1, 3)
        adj, vaf = rand_dag(rng, n, S, rng.choice([0.3, 0.6, 0.9]), rng.choice([0.15, 0.3, 0.5]))
        mt = rng.choice([100000, 7]); mg = rng.choice([100000000, 40])
        s = Search(n, S, adj, vaf.tolist(), 0.1, max_trees=mt, max_grow_calls=mg)
        s.get_lineage_trees(); ev = s.evaluate()
        r = oracle.search(adj, vaf, 0.1, top_s=50, max_trees=mt, max_grow=mg)
        assert r["n_trees"] == len(ev) and r["n_grow"] == s.num_grow
        for k in range(min(50, len(ev))):
            sc, i, t = ev[k]
            assert sc == r["scores"][k] and i == r["dfs_index"][k]
            assert tuple(r["parents"][k]) == parent_vec(t, n) and list(r["orders"][k]) == t.nodes


def test_published_reference_trees(networks):
    """img_demo/RK26.txt.dot.png and img_demo/RMH008.txt.dot.png: the reference's published top
    trees for BASELINE configs[0] and configs[1] -- same cluster sizes, same edges."""
    pub = json.load(open(os.path.join(GOLD, "published_trees.json")))
    for p in pub:
        e = [x for x in networks if (x["dataset"], x["clustering"], x["complete_edges"]) ==
             (p["dataset"], p["clustering"], p["complete_edges"])][0]
        assert {e["tags"][i]: e["sizes"][i] for i in range(1, len(e["tags"]))} == p["nodes"]
        r = oracle.search(e["adj"], np.array(e["aaf"]), e["vaf_error_margin"], top_s=1)
        par = r["parents"][0]
        edges = sorted(["GL" if par[v] == 0 else e["tags"][par[v]], e["tags"][v]] for v in range(1, len(par)))
        assert edges == sorted(p["edges"])


def test_golden_networks_regression(networks):
    for e in networks:
        r = oracle.search(e["adj"], np.array(e["aaf"]), e["vaf_error_margin"], top_s=100)
        g = e["reference_search"]
        assert (r["n_trees"], r["n_grow"], r["n_checks"]) == (g["n_trees"], g["n_grow"], g["n_checks"])
        assert [float(x) for x in r["scores"]] == g["scores"]
        assert [[int(x) for x in p] for p in r["parents"]] == g["parents"]
        assert [[int(x) for x in p] for p in r["orders"]] == g["orders"]


def check_canonical_against_reference(adj, vaf, margin):
    a = oracle.search(adj, vaf, margin, top_s=1, max_trees=10**7, want_all=300000)
    c = oracle.canonical_search(adj, vaf, margin, top_s=1024, max_trees=10**7)
    assert a["n_trees"] == c["n_trees"]
    ref = {bytes(np.asarray(p, dtype=np.int32)): s for p, s in zip(a["all_parents"], a["all_scores"])}
    assert len(ref) == a["n_trees"]                      # the reference never emits a tree twice
    rs = np.sort(a["all_scores"])
    for k in range(len(c["scores"])):
        key = bytes(np.asarray(c["parents"][k], dtype=np.int32))
        assert key in ref
        assert abs(ref[key] - c["scores"][k]) <= 1e-9 * max(abs(c["scores"][k]), 1e-300)
        assert abs(rs[k] - c["scores"][k]) <= 1e-9 * max(abs(c["scores"][k]), 1e-300)
    assert list(c["seq"]) == sorted(c["seq"]) or len(set(c["scores"])) > 1


def test_canonical_order_finds_the_reference_tree_set(networks):
    """The device's canonical enumeration (restated sequentially) finds exactly the reference's
    valid trees with the reference's scores (1e-9); only the order among equal scores differs."""
    rng = random.Random(11)
    for _ in range(60):
        n = rng.randint(3, 9); S = rng.randint(1, 4)
        adj, vaf = rand_dag(rng, n, S, rng.choice([0.3, 0.6, 0.9]), rng.choice([0.15, 0.3, 0.5]))
        check_canonical_against_reference(adj, vaf, 0.1)
    for e in networks:
        if e["reference_search"]["n_trees"] < 100000:
            check_canonical_against_reference(e["adj"], np.array(e["aaf"]), e["vaf_error_margin"])


def test_canonical_tree_cap_is_a_prefix():
    rng = random.Random(5)
    adj, vaf = rand_dag(rng, 9, 2, 0.9, 0.1)
    full = oracle.canonical_search(adj, vaf, 0.1, top_s=1000, max_trees=10**7)
    assert full["n_trees"] > 500
    cut = oracle.canonical_search(adj, vaf, 0.1, top_s=1000, max_trees=200)
    assert cut["n_trees"] == 200 and all(s < 200 for s in cut["seq"])


def test_synthetic_generator_is_deterministic_and_a_dag():
    from lichee_b200.synth import make_instance
    a, _ = make_instance(20, 8, seed=7)
    b, _ = make_instance(20, 8, seed=7)
    assert a.adj == b.adj and a.aaf == b.aaf
    indeg = [0] * a.n
    for ch in a.adj:
        for c in ch:
            indeg[c] += 1
    assert indeg[0] == 0 and all(d > 0 for d in indeg[1:])


def test_reference_enumeration_order_depends_on_search_history():
    """Why the device cannot reproduce the reference's order among equal scores:
    the literal restatement (mode 0: ArrayList re-appends of PHYNetwork.addEdge / the frontier stack)
    and the same algorithm with side-effect-free nested calls (mode 1) find the SAME trees but in a
    different order on dense networks, i.e. the order is a function of the whole search history, not
    of the tree or its path."""
    from lichee_b200.synth import make_instance
    diverged = 0
    for seed in range(6):
        net, prm = make_instance(10, 4, seed)
        a = oracle.search(net.adj, np.array(net.aaf), prm.vaf_error_margin, top_s=1, max_trees=20000, want_all=20000)
        b = oracle.search(net.adj, np.array(net.aaf), prm.vaf_error_margin, top_s=1, max_trees=20000, want_all=20000, mode=1)
        if a["n_trees"] < 20000:      # both ran to completion: same set of trees
            assert sorted(map(bytes, a["all_parents"])) == sorted(map(bytes, b["all_parents"]))
        k = min(len(a["all_parents"]), len(b["all_parents"]))
        if (a["all_parents"][:k] != b["all_parents"][:k]).any():
            diverged += 1
    assert diverged >= 3


def test_partial_tree_score_is_a_lower_bound():
    """What the device's pruning rests on: with non-negative VAFs the canonical score
    of a tree WITHOUT its last one or two nodes (same fixed-shape reduction, the missing rows contribute
    exact zeros) never exceeds the score of the complete tree -- in floating point, not only in real
    arithmetic.  Checked on the ranked trees of random synthetic networks."""
    from lichee_b200.synth import make_instance
    checked = 0
    for seed in range(8):
        net, prm = make_instance(9 + seed, 3 + 4 * (seed % 3), seed=100 + seed, noise=0.02)
        vaf = np.array(net.aaf, dtype=np.float64)
        assert (vaf >= 0).all()
        ref = oracle.canonical_search(net.adj, vaf, prm.vaf_error_margin, top_s=300, max_trees=200000)
        sigma = [int(x) for x in ref["sigma"]]
        for k in range(len(ref["scores"])):
            parent = [int(x) for x in ref["parents"][k]]
            ok, full = oracle.canonical_eval_tree(vaf, prm.vaf_error_margin, parent)
            assert ok and full == ref["scores"][k]
            for drop in (1, 2):
                gone = set(sigma[-drop:])
                keep = [v for v in range(len(parent)) if v not in gone]
                if any(parent[v] in gone for v in keep):
                    continue                      # a kept node hangs under a dropped one: not a partial tree
                new_id = {v: i for i, v in enumerate(keep)}
                sub_parent = [-1 if parent[v] < 0 else new_id[parent[v]] for v in keep]
                ok2, part = oracle.canonical_eval_tree(vaf[keep], prm.vaf_error_margin, sub_parent)
                assert ok2, "a sub-tree of a valid tree is valid"
                assert part <= full, (seed, k, drop, part, full)
                checked += 1
    assert checked > 500
