"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Two layers, both from oracle/lichee_oracle.c:
  * canonical_search  -- sequential statement of the device's enumeration/summation order:
                         compared BIT FOR BIT (tree count, ranked trees, scores, seq numbers).
  * search (mode 0)   -- the reference's own Gabow-Myers search, restated literally: compared on
                         everything that does not depend on its history-dependent enumeration
                         order: the SET of valid trees, each tree's score (<= 1e-9 relative, the
                         tolerance BASELINE.json states), the ranked score list, and the ranked
                         trees themselves whenever scores are distinct.
"""
import random

import numpy as np
import pytest

import oracle
from lichee_b200.synth import make_instance

pytestmark = pytest.mark.gpu

REL = 1e-9


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


def gpu_search(adj, vaf, margin, num_save, **kw):
    """Canonical-order search unless the caller passes ref_order_budget (tests/test_gpu_reference_order.py
    covers the reference-order mode)."""
    import lichee_b200
    kw.setdefault("ref_order_budget", -1)
    s = lichee_b200.LineageSearch(adj, vaf, margin, num_save=num_save, **kw)
    try:
        st = s.run()
        return st.as_dict(), s.trees(), list(s.sigma())
    finally:
        s.close()


def assert_matches_canonical(adj, vaf, margin, num_save, max_trees=10**9, **kw):
    st, trees, sigma = gpu_search(adj, vaf, margin, num_save, max_num_trees=max_trees, **kw)
    ref = oracle.canonical_search(adj, vaf, margin, top_s=num_save, max_trees=max_trees)
    assert sigma == list(ref["sigma"])
    assert st["num_trees"] == ref["n_trees"]
    assert len(trees) == len(ref["scores"])
    for k, t in enumerate(trees):
        assert t.errorScore == ref["scores"][k], (k, t.errorScore, ref["scores"][k])   # bit-exact double
        assert t.parent == list(ref["parents"][k]), k
        assert t.seq == ref["seq"][k], k
    if not (st["status"] & 3):
        assert st["num_checks"] == ref["n_checks"]
        assert st["num_grow_calls"] == ref["n_grow"]
    return st, trees, ref


@pytest.mark.parametrize("seed", range(12))
def test_random_dags_bit_exact_vs_canonical(built, seed):
    rng = random.Random(seed)
    n = rng.randint(2, 11); S = rng.randint(1, 6)
    adj, vaf = rand_dag(rng, n, S, rng.choice([0.3, 0.6, 0.9]), rng.choice([0.15, 0.3, 0.5]))
    assert_matches_canonical(adj, vaf, 0.1, num_save=rng.choice([1, 7, 100]))


@pytest.mark.parametrize("seed", range(6))
def test_synthetic_complete_network_vs_canonical(built, seed):
    net, prm = make_instance(11, 5, seed)
    st, trees, ref = assert_matches_canonical(net.adj, net.aaf, prm.vaf_error_margin, num_save=100, max_trees=300000)
    assert st["num_trees"] > 0


def test_more_than_32_samples_two_per_lane(built):
    net, prm = make_instance(12, 40, seed=5)
    assert_matches_canonical(net.adj, net.aaf, prm.vaf_error_margin, num_save=50, max_trees=200000)
    net, prm = make_instance(14, 64, seed=6)
    assert_matches_canonical(net.adj, net.aaf, prm.vaf_error_margin, num_save=1000, max_trees=100000)


@pytest.mark.parametrize("n_clusters,n_samples", [(17, 7), (20, 20), (33, 12), (40, 20), (70, 33), (100, 48), (120, 64), (127, 64)])
def test_every_item_width_and_the_named_shapes(built, n_clusters, n_samples):
    """Item strides 16..128 bytes (the byte layout of a frontier item depends on it), including the
    BASELINE shapes 40x20 and 120x64 and the size limits; capped so the CPU specification stays fast."""
    net, prm = make_instance(n_clusters, n_samples, seed=n_clusters)
    assert_matches_canonical(net.adj, net.aaf, prm.vaf_error_margin, num_save=200, max_trees=30000)


def test_chunked_frontier_small_level_capacity(built):
    """A tiny level_capacity forces many depth-first chunks and partial emits; results must not move."""
    net, prm = make_instance(11, 4, seed=2)
    a = assert_matches_canonical(net.adj, net.aaf, prm.vaf_error_margin, num_save=64, max_trees=10**9)[0]
    b = assert_matches_canonical(net.adj, net.aaf, prm.vaf_error_margin, num_save=64, max_trees=10**9,
                                 level_capacity=1024)[0]
    assert a["num_trees"] == b["num_trees"] and b["num_launches"] > a["num_launches"]


def test_tree_cap_is_a_prefix_of_the_canonical_order(built):
    """Parameters.MAX_NUM_TREES: the first max_trees valid trees in enumeration order are kept."""
    net, prm = make_instance(12, 4, seed=1)
    for cap in (1, 17, 1000, 100000):
        st, trees, ref = assert_matches_canonical(net.adj, net.aaf, prm.vaf_error_margin, num_save=10, max_trees=cap)
        assert st["num_trees"] <= cap
        if st["num_trees"] == cap:
            assert st["status"] & 1
            assert all(t.seq < cap for t in trees)


@pytest.mark.parametrize("seed", range(10))
def test_valid_tree_set_and_scores_vs_reference_search(built, seed):
    """Against the literal restatement of PHYNetwork.grow: same set of valid trees, same scores
    (1e-9 relative), same ranked score list; identical ranked trees where scores are distinct."""
    rng = random.Random(100 + seed)
    if seed % 2:
        net, prm = make_instance(8, 3, seed)
        adj, vaf, margin = net.adj, np.array(net.aaf), prm.vaf_error_margin
    else:
        n = rng.randint(3, 9); S = rng.randint(1, 4)
        adj, vaf = rand_dag(rng, n, S, rng.choice([0.4, 0.8]), rng.choice([0.2, 0.4]))
        margin = 0.1
    ref = oracle.search(adj, vaf, margin, top_s=1024, max_trees=10**7, want_all=10**6)
    st, trees, _ = gpu_search(adj, vaf, margin, 1024, max_num_trees=10**9)
    assert st["num_trees"] == ref["n_trees"]
    ref_by_tree = {bytes(np.asarray(p, dtype=np.int32)): s for p, s in zip(ref["all_parents"], ref["all_scores"])}
    assert len(ref_by_tree) == ref["n_trees"]
    for k, t in enumerate(trees):
        key = bytes(np.asarray(t.parent, dtype=np.int32))
        assert key in ref_by_tree
        assert abs(ref_by_tree[key] - t.errorScore) <= REL * max(abs(t.errorScore), 1e-300)
        assert abs(ref["scores"][k] - t.errorScore) <= REL * max(abs(t.errorScore), 1e-300)
    # where neighbouring reference scores differ by more than the tolerance the ranked trees are identical
    rs = ref["scores"]
    for k, t in enumerate(trees):
        lo = k == 0 or rs[k] - rs[k - 1] > 1e-7 * max(rs[k], 1e-12)
        hi = k + 1 >= len(rs) or rs[k + 1] - rs[k] > 1e-7 * max(rs[k], 1e-12)
        if lo and hi and len(trees) == st["num_trees"]:
            assert t.parent == list(ref["parents"][k])


def test_edge_cases(built):
    import lichee_b200
    # root only: getLineageTrees returns an empty list (PHYNetwork.java:557)
    st, trees, _ = gpu_search([[]], np.array([[0.5]]), 0.1, 1)
    assert st["num_trees"] == 0 and trees == []
    # a node no edge reaches: no spanning tree
    st, trees, _ = gpu_search([[1], [], []], np.array([[0.5], [0.2], [0.1]]), 0.1, 1)
    assert st["num_trees"] == 0
    # single chain
    st, trees, _ = gpu_search([[1], [2], []], np.array([[0.5], [0.3], [0.2]]), 0.1, 5)
    assert st["num_trees"] == 1 and trees[0].parent == [-1, 0, 1] and trees[0].errorScore == 0.0
    # every tree violates the sum rule: zero valid trees
    st, trees, _ = gpu_search([[1, 2], [], []], np.array([[0.5], [0.45], [0.45]]), 0.1, 5)
    assert st["num_trees"] == 0
    # duplicate edges collapse (PHYNetwork.addEdge refuses duplicates)
    st, trees, _ = gpu_search([[1, 1, 2], [2], []], np.array([[0.5], [0.3], [0.2]]), 0.1, 5)
    assert st["num_trees"] == 2
    # a directed cycle: the canonical parent-choice search needs a DAG (the reference-order replay does not,
    # tests/test_gpu_reference_order.py)
    with pytest.raises(lichee_b200.LicheeError) as e:
        gpu_search([[1], [2], [1]], np.array([[0.5], [0.3], [0.2]]), 0.1, 1)
    assert e.value.code == -3
    with pytest.raises(lichee_b200.LicheeError):
        gpu_search([[1], []], np.zeros((2, 65)), 0.1, 1)


def test_sharded_search_merges_to_the_single_gpu_answer(built):
    """part_rank/part_count slices of the canonical order, merged like the NCCL allgather does."""
    import lichee_b200
    net, prm = make_instance(11, 5, seed=4)
    whole, trees, _ = gpu_search(net.adj, net.aaf, prm.vaf_error_margin, 32, max_num_trees=10**9)
    for P in (2, 3, 8):
        parts = [lichee_b200.LineageSearch(net.adj, net.aaf, prm.vaf_error_margin, num_save=32, max_num_trees=10**9,
                                           part_rank=r, part_count=P) for r in range(P)]
        stats = [p.run() for p in parts]
        assert sum(s.num_trees for s in stats) == whole["num_trees"]
        recs = np.concatenate([p.export_ranked() for p in parts])
        parts[0].merge_ranked(recs, P)
        merged = parts[0].trees()
        assert [t.parent for t in merged] == [t.parent for t in trees]
        assert [t.errorScore for t in merged] == [t.errorScore for t in trees]
        for p in parts:
            p.close()


def test_merge_of_long_sorted_lists(built):
    """More than one rank-merge tile of records (3 GPUs x 700 trees): the merge form of the selection (each list
    arrives sorted; bitonic merge per list) must give the single-search list, ties included."""
    import lichee_b200
    e = [x for x in _networks() if x["dataset"] == "ccRCC/RK26.txt" and x["complete_edges"] and x["clustering"] == "single"][0]
    for adj, vaf, margin in ((e["adj"], np.array(e["aaf"]), e["vaf_error_margin"]),) + tuple(
            (n.adj, np.array(n.aaf), p.vaf_error_margin) for n, p in (make_instance(11, 5, seed=4),)):
        whole, trees, _ = gpu_search(adj, vaf, margin, 700, max_num_trees=10**9)
        P = 3
        parts = [lichee_b200.LineageSearch(adj, vaf, margin, num_save=700, max_num_trees=10**9, part_rank=r, part_count=P)
                 for r in range(P)]
        for q in parts:
            q.run()
        recs = np.concatenate([q.export_ranked() for q in parts])
        parts[1].merge_ranked(recs, P)
        merged = parts[1].trees()
        assert [(t.errorScore, t.parent) for t in merged] == [(t.errorScore, t.parent) for t in trees]
        for q in parts:
            q.close()


def _networks():
    import json, os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "networks_bundled.json")))


def test_bundled_datasets_match_the_reference_search(built):
    """BASELINE configs[0..2]: every bundled ccRCC / HGSC network (README settings, with and
    without -c) through the CUDA path against what the literal reference search returned when the
    fixture was generated: same number of valid trees, same ranked scores (1e-9), and the same
    ranked trees wherever the reference's scores are distinct."""
    for e in _networks():
        g = e["reference_search"]
        if g["n_trees"] >= 100000:          # the reference stopped at MAX_NUM_TREES: a DFS prefix, not comparable
            continue
        st, trees, _ = gpu_search(e["adj"], np.array(e["aaf"]), e["vaf_error_margin"], 100, max_num_trees=10**9)
        assert st["num_trees"] == g["n_trees"], e["dataset"]
        assert len(trees) == len(g["scores"])
        for k, t in enumerate(trees):
            assert abs(g["scores"][k] - t.errorScore) <= REL * max(abs(t.errorScore), 1e-300)
        rs = g["scores"]
        for k, t in enumerate(trees):
            lo = k == 0 or rs[k] - rs[k - 1] > 1e-7 * max(rs[k], 1e-12)
            hi = k + 1 >= len(rs) or rs[k + 1] - rs[k] > 1e-7 * max(rs[k], 1e-12)
            if lo and hi and len(trees) == g["n_trees"]:
                assert t.parent == g["parents"][k], (e["dataset"], k)


def test_published_reference_trees_on_gpu(built):
    """The top trees the reference publishes for ccRCC RK26 and RMH008 (img_demo/*.dot.png)."""
    import json, os
    pub = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "published_trees.json")))
    nets = _networks()
    for p in pub:
        e = [x for x in nets if (x["dataset"], x["clustering"], x["complete_edges"]) ==
             (p["dataset"], p["clustering"], p["complete_edges"])][0]
        st, trees, _ = gpu_search(e["adj"], np.array(e["aaf"]), e["vaf_error_margin"], 1)
        par = trees[0].parent
        edges = sorted(["GL" if par[v] == 0 else e["tags"][par[v]], e["tags"][v]] for v in range(1, len(par)))
        assert edges == sorted(p["edges"])
        # treeNodes / treeEdges reconstruction is consistent
        assert trees[0].treeNodes[0] == 0 and sorted(trees[0].treeNodes) == list(range(len(par)))
        assert all(par[c] == q for q, ch in trees[0].treeEdges.items() for c in ch)


def test_work_stealing_slices_give_the_whole_search(built):
    """Slices claimed one after another on one handle (ranked list and threshold kept), and the same
    slices dealt to two handles out of order and merged: both equal the single whole search."""
    import lichee_b200
    net, prm = make_instance(11, 5, seed=4)
    whole, trees, _ = gpu_search(net.adj, net.aaf, prm.vaf_error_margin, 40, max_num_trees=10**9)
    T = 12
    s = lichee_b200.LineageSearch(net.adj, net.aaf, prm.vaf_error_margin, num_save=40, max_num_trees=10**9, ref_order_budget=-1)
    todo = iter(range(T + 1))
    assert s.run_claimed(T, lambda: next(todo))
    st = s.stats()
    assert st.num_trees == whole["num_trees"] and st.num_scored == whole["num_trees"]
    got = s.trees()
    assert [t.parent for t in got] == [t.parent for t in trees]
    assert [t.errorScore for t in got] == [t.errorScore for t in trees]
    s.close()
    a = lichee_b200.LineageSearch(net.adj, net.aaf, prm.vaf_error_margin, num_save=40, max_num_trees=10**9, ref_order_budget=-1)
    b = lichee_b200.LineageSearch(net.adj, net.aaf, prm.vaf_error_margin, num_save=40, max_num_trees=10**9, ref_order_budget=-1)
    c = lichee_b200.LineageSearch(net.adj, net.aaf, prm.vaf_error_margin, num_save=40, max_num_trees=10**9, ref_order_budget=-1)
    order_a, order_b = iter([0, 2, 5, 7, 11, T]), iter([1, 3, 4, 6, 8, 9, 10, T])       # (ascending on each handle)
    a.run_claimed(T, lambda: next(order_a))
    b.run_claimed(T, lambda: next(order_b))
    assert not c.run_claimed(T, lambda: T)                  # a rank that found nothing left to claim
    assert a.stats().num_trees + b.stats().num_trees == whole["num_trees"] and c.stats().num_trees == 0
    recs = np.concatenate([a.export_ranked(), b.export_ranked(), c.export_ranked()])
    for h in (a, b, c):
        h.merge_ranked(recs, 3)
        m = h.trees()
        assert [t.parent for t in m] == [t.parent for t in trees]
        assert [t.errorScore for t in m] == [t.errorScore for t in trees]
        h.close()


def test_slices_with_tied_scores_and_the_ascending_order_rule(built):
    """Real -c data is dominated by zero-error ties (RK26 -c: 99 of the top 100 scores tie).  Ties rank by (slice,
    seq); a kept list admits a tying tree only if it precedes the held ones, so slices kept on one handle must be
    searched in ascending order -- anything else is refused."""
    import lichee_b200
    e = [x for x in _networks() if x["dataset"] == "ccRCC/RK26.txt" and x["complete_edges"] and x["clustering"] == "single"][0]
    adj, vaf, margin = e["adj"], np.array(e["aaf"]), e["vaf_error_margin"]
    whole, trees, _ = gpu_search(adj, vaf, margin, 100, max_num_trees=10**9)
    assert len({t.errorScore for t in trees}) < 10                      # massive ties
    T = 9
    a = lichee_b200.LineageSearch(adj, vaf, margin, num_save=100, max_num_trees=10**9, ref_order_budget=-1)
    b = lichee_b200.LineageSearch(adj, vaf, margin, num_save=100, max_num_trees=10**9, ref_order_budget=-1)
    ia, ib = iter([0, 3, 4, 8, T]), iter([1, 2, 5, 6, 7, T])
    a.run_claimed(T, lambda: next(ia))
    b.run_claimed(T, lambda: next(ib))
    assert a.stats().num_trees + b.stats().num_trees == whole["num_trees"]
    recs = np.concatenate([a.export_ranked(), b.export_ranked()])
    a.merge_ranked(recs, 2)
    assert [(t.errorScore, t.parent) for t in a.trees()] == [(t.errorScore, t.parent) for t in trees]
    with pytest.raises(lichee_b200.LicheeError) as err:                 # slice 2 after slice 7 on the same list
        b.run_slice(2, T, keep_ranked=True)
    assert err.value.code == -6
    a.close(); b.close()


@pytest.mark.parametrize("n_clusters,n_samples,cap", [(40, 20, 1 << 22), (120, 64, 1 << 22)])
def test_full_size_properties(built, n_clusters, n_samples, cap):
    """BASELINE configs[3] and [4] at a tree count the CPU cannot enumerate: properties that do not
    depend on size.  Every returned tree is a spanning tree that passes the sum rule and carries
    exactly the score the CPU evaluator computes for it; the list is ranked by (score, seq); the
    answer does not move with the frontier capacity (chunking), with a re-run on the same handle, or
    with the choice of slices; and a prefix cap only ever removes trees (seq < cap)."""
    import lichee_b200
    net, prm = make_instance(n_clusters, n_samples, seed=2026)
    vaf = np.array(net.aaf)
    s = lichee_b200.LineageSearch(net.adj, vaf, prm.vaf_error_margin, num_save=1000, max_num_trees=cap, ref_order_budget=-1)
    st = s.run()
    trees = s.trees()
    assert st.num_trees == cap and st.num_scored == cap and (st.status & 1)
    assert len(trees) == 1000
    edges = {(p, c) for p, ch in enumerate(net.adj) for c in ch}
    for t in trees:
        assert all((t.parent[v], v) in edges for v in range(1, net.n))       # uses network edges only
        assert sorted(t.treeNodes) == list(range(net.n))                      # spans every node
        ok, sc = oracle.canonical_eval_tree(vaf, prm.vaf_error_margin, t.parent)
        assert ok and sc == t.errorScore                                      # valid, bit-equal score
        assert 0 <= t.seq < cap
    keys = [(t.errorScore, t.seq) for t in trees]
    assert keys == sorted(keys) and len(set(keys)) == len(keys)
    again = s.run()
    assert again.num_trees == st.num_trees and [t.parent for t in s.trees()] == [t.parent for t in trees]   # idempotent
    s.close()
    small = lichee_b200.LineageSearch(net.adj, vaf, prm.vaf_error_margin, num_save=1000, max_num_trees=cap,
                                      level_capacity=1 << 16, ref_order_budget=-1)
    st2 = small.run()
    t2 = small.trees()
    assert st2.num_trees == cap and st2.num_launches > st.num_launches
    assert [(t.errorScore, t.seq, t.parent) for t in t2] == [(t.errorScore, t.seq, t.parent) for t in trees]
    small.close()
    # a tighter prefix cap keeps exactly the trees of the longer run whose seq is below it
    half = lichee_b200.LineageSearch(net.adj, vaf, prm.vaf_error_margin, num_save=1000, max_num_trees=cap // 4, ref_order_budget=-1)
    half.run()
    th = half.trees()
    best_possible = sorted((t.errorScore, t.seq) for t in trees if t.seq < cap // 4)
    assert all(t.seq < cap // 4 for t in th)
    kept = {(t.errorScore, t.seq) for t in th}
    assert set(best_possible) <= kept
    half.close()


def test_grow_call_cap_and_memory_release(built):
    """Parameters.MAX_NUM_GROW_CALLS stops the search at the first chunk boundary after the count of
    accepted partial trees reaches it (status bit 2); the pooled buffers can be handed back."""
    import lichee_b200
    net, prm = make_instance(14, 6, seed=9)
    st, trees, _ = gpu_search(net.adj, net.aaf, prm.vaf_error_margin, 10, max_num_trees=10**12, max_num_grow_calls=5000)
    assert st["status"] & 2 and st["num_grow_calls"] >= 5000
    full, _, _ = gpu_search(net.adj, net.aaf, prm.vaf_error_margin, 10, max_num_trees=10**12)
    assert not (full["status"] & 3) and full["num_grow_calls"] > st["num_grow_calls"]
    lichee_b200.release_cached_memory()
    again, _, _ = gpu_search(net.adj, net.aaf, prm.vaf_error_margin, 10, max_num_trees=10**12)
    assert again["num_trees"] == full["num_trees"]


def test_randomized_stress_bit_exact_vs_canonical(built):
    """150 seeded instances over node count, sample count (1..64), density, VAF scale, margin, -s,
    tree cap and frontier capacity: CUDA == sequential canonical specification, bit for bit."""
    rng = random.Random(20261019)
    done = 0
    for it in range(150):
        n = rng.randint(2, 26)
        S = rng.choice([1, 2, 3, 5, 8, 13, 20, 31, 32, 33, 47, 64])
        adj, vaf = rand_dag(rng, n, S, rng.choice([0.15, 0.3, 0.6, 0.9]), rng.choice([0.05, 0.15, 0.3, 0.5]))
        margin = rng.choice([0.0, 0.01, 0.1, 0.25])
        cap = rng.choice([1, 3, 100, 5000, 40000])
        kw = {}
        if rng.random() < 0.4:
            kw["level_capacity"] = rng.choice([1024, 2048, 8192])
        assert_matches_canonical(adj, vaf, margin, num_save=rng.choice([1, 2, 33, 1000]), max_trees=cap, **kw)
        done += 1
    assert done == 150


@pytest.mark.parametrize("seed", range(4))
def test_mid_size_tree_set_vs_reference_search(built, seed):
    """Sparser networks of 14-18 nodes on which the literal reference search still runs to completion
    (10^4-10^6 trees): identical number of valid trees, identical ranked scores (1e-9)."""
    rng = random.Random(500 + seed)
    n = rng.randint(14, 18); S = rng.randint(2, 6)
    adj, vaf = rand_dag(rng, n, S, 0.3, 0.05)
    ref = oracle.search(adj, vaf, 0.1, top_s=1000, max_trees=3 * 10**6, max_grow=5 * 10**7, want_all=3 * 10**6)
    if ref["stopped"]:
        pytest.skip("reference search hit a cap on this instance")
    st, trees, _ = gpu_search(adj, vaf, 0.1, 1000, max_num_trees=10**12)
    assert st["num_trees"] == ref["n_trees"] > 5000
    assert len(trees) == len(ref["scores"])
    by_tree = {bytes(np.asarray(p, dtype=np.int32)): sc for p, sc in zip(ref["all_parents"], ref["all_scores"])}
    for k, t in enumerate(trees):
        assert abs(ref["scores"][k] - t.errorScore) <= REL * max(abs(t.errorScore), 1e-300)
        key = bytes(np.asarray(t.parent, dtype=np.int32))
        assert key in by_tree and abs(by_tree[key] - t.errorScore) <= REL * max(abs(t.errorScore), 1e-300)


def test_scheduler_and_bound_do_not_change_results(built, monkeypatch):
    """The single-block scheduler of narrow chunks (descend_kernel) and the lower bound that lets the last
    two levels count whole sibling groups without scoring them can be switched off (LICHEE_NO_NARROW,
    LICHEE_NO_PRUNE): tree counts, check counts, ranked trees, scores and sequence numbers must not move,
    and the bound must actually have skipped work when it is on."""
    cases = [make_instance(12, 5, seed=4), make_instance(30, 12, seed=7), make_instance(14, 40, seed=8)]
    for net, prm in cases:
        base = None
        for env in ({}, {"LICHEE_NO_NARROW": "1"}, {"LICHEE_NO_PRUNE": "1"},
                    {"LICHEE_NO_NARROW": "1", "LICHEE_NO_PRUNE": "1"}):
            for k in ("LICHEE_NO_NARROW", "LICHEE_NO_PRUNE"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            st, trees, _ = gpu_search(net.adj, net.aaf, prm.vaf_error_margin, 25, max_num_trees=2_000_000)
            key = (st["num_trees"], [(t.errorScore, t.seq, t.parent) for t in trees])
            if base is None:
                base = key
            else:
                assert key == base, env
    for k in ("LICHEE_NO_NARROW", "LICHEE_NO_PRUNE"):
        monkeypatch.delenv(k, raising=False)


def test_two_node_and_three_node_networks(built):
    """The last two tree levels have their own kernels (check2 / emit2 / leaf): the smallest networks, where
    those levels are the first ones, against the specification."""
    # root + one node; root + two nodes in every edge configuration that keeps the root the only source
    assert_matches_canonical([[1], []], np.array([[0.5, 0.5], [0.2, 0.3]]), 0.1, num_save=5)
    vaf = np.array([[0.5, 0.5, 0.5], [0.3, 0.2, 0.25], [0.2, 0.2, 0.1]])
    for adj in ([[1, 2], [], []], [[1, 2], [2], []], [[1], [2], []]):
        assert_matches_canonical(adj, vaf, 0.1, num_save=5)
    # nothing fits under anything but the root
    assert_matches_canonical([[1, 2], [2], []], np.array([[0.5, 0.5], [0.4, 0.4], [0.45, 0.45]]), 0.0, num_save=3)
