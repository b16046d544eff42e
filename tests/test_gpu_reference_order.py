"""Reference-order mode on the GPU (replay_kernel + score_ref_kernel + the selection kernels) against the
LITERAL restatement of PHYNetwork.grow / PHYTree.computeErrorScore / Collections.sort in oracle/ (mode 0):
everything the reference writes must be identical -- number of valid trees, numGrowCalls, the ranked list
(parents, treeNodes order, discovery index) and the scores BIT FOR BIT (same additions in the same order).
This is the default mode of the library (ref_order_budget = 0)."""
import json
import os
import random

import numpy as np
import pytest

import oracle
from lichee_b200.synth import make_instance

pytestmark = pytest.mark.gpu


def ref_search(adj, vaf, margin, num_save, **kw):
    import lichee_b200
    s = lichee_b200.LineageSearch(adj, vaf, margin, num_save=num_save, **kw)
    try:
        st = s.run()
        return st.as_dict(), s.trees()
    finally:
        s.close()


def assert_identical_to_reference(adj, vaf, margin, num_save, max_trees=100000, max_grow=100000000, **kw):
    st, trees = ref_search(adj, vaf, margin, num_save, max_num_trees=max_trees, max_num_grow_calls=max_grow, **kw)
    ref = oracle.search(adj, vaf, margin, top_s=num_save, max_trees=max_trees, max_grow=max_grow)
    assert not (st["status"] & 4), "fell back to the canonical order"
    assert st["num_trees"] == ref["n_trees"]
    assert len(trees) == len(ref["scores"])
    for k, t in enumerate(trees):
        assert t.errorScore == ref["scores"][k], (k, t.errorScore, ref["scores"][k])      # bit-exact double
        assert t.parent == list(ref["parents"][k]), k
        assert t.treeNodes == list(ref["orders"][k]), k                                   # PHYTree.treeNodes
        assert t.seq == ref["dfs_index"][k], k                                            # position in spanningTrees before the sort
    if not ref["stopped"]:
        # (after a cap the reference keeps unwinding and counting; the counters are only comparable without one)
        assert st["num_grow_calls"] == ref["n_grow"]
        assert st["num_checks"] == ref["n_checks"]
        assert not (st["status"] & 3)
    else:
        assert st["status"] & 3
    return st, trees, ref


def _networks():
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "networks_bundled.json")))


@pytest.mark.parametrize("num_save", [1, 100])
def test_all_bundled_networks_identical_to_the_reference(built, num_save):
    """BASELINE configs[0..2]: all 56 ccRCC / HGSC networks (README settings, with and without -c, two
    clusterings) with the reference's default caps: the saved top-s list is the reference's, tie order
    included (with the canonical order the top tree differed on 10 of the 51 networks that
    have trees and the top-100 list on 20; tests/test_oracle_cpu.py pins those counts)."""
    for e in _networks():
        adj, vaf = e["adj"], np.array(e["aaf"])
        st, trees, ref = assert_identical_to_reference(adj, vaf, e["vaf_error_margin"], num_save)
        # and against what the literal port returned when the fixture was generated (committed)
        g = e["reference_search"]
        assert st["num_trees"] == g["n_trees"] and st["num_grow_calls"] == g["n_grow"], e["dataset"]
        m = min(num_save, len(g["scores"]))
        assert [t.errorScore for t in trees[:m]] == g["scores"][:m], e["dataset"]
        assert [t.parent for t in trees[:m]] == g["parents"][:m], e["dataset"]


def rand_graph(rng, n, S, p_edge, vaf_scale, cyclic):
    adj = [[] for _ in range(n)]
    order = list(range(1, n)); rng.shuffle(order); order = [0] + order
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            if j > i or (cyclic and rng.random() < 0.3):
                if rng.random() < p_edge:
                    adj[order[i]].append(order[j])
    for a in adj:
        rng.shuffle(a)
    vaf = np.array([[0.5] * S] + [[rng.random() * vaf_scale for _ in range(S)] for _ in range(n - 1)])
    return adj, vaf


def test_random_digraphs_caps_and_cycles(built):
    """250 seeded digraphs (a third of them cyclic, some with edges into the root, self loops and duplicate
    edges), 1..64 samples, both caps at small values: identical to the literal port."""
    rng = random.Random(20261020)
    for it in range(250):
        n = rng.randint(1, 10)
        S = rng.choice([1, 2, 3, 5, 8, 20, 31, 32, 33, 47, 64])
        adj, vaf = rand_graph(rng, n, S, rng.choice([0.3, 0.6, 0.9]), rng.choice([0.1, 0.2, 0.4]), rng.random() < 0.33)
        if n > 2 and rng.random() < 0.2:
            adj[rng.randrange(1, n)].append(0)                 # edge into the root
            u = rng.randrange(n); adj[u].append(u)             # self loop
            if adj[0]:
                adj[0].append(adj[0][0])                       # duplicate
        # (the literal port keeps self loops / root in-edges; neither is ever followed, but a self loop keeps the
        #  reference's bridge test from ending a loop early -- exclude those instances from the comparison)
        clean = [[v for k, v in enumerate(a) if v != u and v != 0 and v not in a[:k]] for u, a in enumerate(adj)]
        assert_identical_to_reference(adj if clean == adj else clean, vaf, rng.choice([0.0, 0.05, 0.1]),
                                      num_save=rng.choice([1, 7, 100, 1000]),
                                      max_trees=rng.choice([1, 5, 1000, 100000]), max_grow=rng.choice([3, 50, 5000, 10**8]))
        if clean != adj:
            a, _ = ref_search(adj, vaf, 0.1, 5)
            b, _ = ref_search(clean, vaf, 0.1, 5)
            assert a["num_trees"] == b["num_trees"] and a["num_grow_calls"] == b["num_grow_calls"]


def test_random_digraphs_33_to_64_nodes(built):
    """Sparse digraphs with 33..64 nodes (more than one 32-node group per list scan of the replay, node lists that do not
    fill the last group), a quarter of them cyclic, searched to the end or into a grow-call cap: identical to the literal
    port."""
    rng = random.Random(64)
    done = 0
    for it in range(40):
        n = rng.randint(33, 64)
        S = rng.choice([1, 4, 20, 33, 64])
        adj, vaf = rand_graph(rng, n, S, rng.choice([0.04, 0.06, 0.09]), rng.choice([0.02, 0.05]), rng.random() < 0.25)
        for v in range(1, n):                                   # every node reachable: an edge from the root if it has none
            if not any(v in a for a in adj):
                adj[0].append(v)
        st, trees, ref = assert_identical_to_reference(adj, vaf, rng.choice([0.05, 0.1]), num_save=rng.choice([1, 100]),
                                                       max_trees=rng.choice([1000, 100000]), max_grow=200000, ref_order_budget=1 << 20)
        done += 0 if ref["stopped"] else 1
    assert done >= 5


def test_many_trees_several_batches_and_ties(built):
    """More trees than one record batch holds (the replay pauses, the batch is scored and merged, the replay
    resumes), -s 1000, zero-margin ties: RK26 -c (28080 trees, 99 of the top 100 scores tie) and synthetic
    networks with 5*10^4 trees under the reference's cap."""
    e = [x for x in _networks() if x["dataset"] == "ccRCC/RK26.txt" and x["complete_edges"] and x["clustering"] == "single"][0]
    st, trees, ref = assert_identical_to_reference(e["adj"], np.array(e["aaf"]), e["vaf_error_margin"], 1000)
    assert st["num_trees"] == 28080 and st["num_launches"] > 3
    for (c, s, seed, cap, hits_cap) in ((11, 5, 3, 50000, True), (12, 4, 1, 100000, True), (12, 40, 5, 60000, False)):
        net, prm = make_instance(c, s, seed)
        st, trees, ref = assert_identical_to_reference(net.adj, np.array(net.aaf), prm.vaf_error_margin, 1000, max_trees=cap,
                                                       ref_order_budget=1 << 24)
        assert (st["num_trees"] == cap and bool(st["status"] & 1)) == hits_cap


def test_budget_exhausted_falls_back_to_the_canonical_order(built):
    """The replay is budgeted in grow() calls; when the reference's own search would take longer the library
    answers in the canonical order and says so (status bit 4).  40 clusters x 20 samples: the reference
    finds no tree in 10^8 calls (tests/golden/reference_caps.json), the canonical search finds 10^5 at once."""
    net, prm = make_instance(40, 20, seed=2026)
    vaf = np.array(net.aaf)
    st, trees = ref_search(net.adj, vaf, prm.vaf_error_margin, 10, ref_order_budget=20000)
    assert st["status"] & 4 and st["status"] & 1 and st["num_trees"] == 100000
    can = oracle.canonical_search(net.adj, vaf, prm.vaf_error_margin, top_s=10, max_trees=100000)
    assert [t.parent for t in trees] == [list(p) for p in can["parents"]]
    assert [t.errorScore for t in trees] == list(can["scores"])
    # with a budget beyond the reference's own MAX_NUM_GROW_CALLS the replay ends like the reference does
    st, trees = ref_search(net.adj, vaf, prm.vaf_error_margin, 10, max_num_grow_calls=30000, ref_order_budget=10**6)
    assert not (st["status"] & 4) and st["status"] & 2 and st["num_trees"] == 0 and trees == []


def test_cyclic_constraint_network(built):
    """Three clusters of one group whose pairwise centroid differences stay inside -e form a directed 3-cycle
    (PHYNetwork.java:114-118,185-232; VERDICT r1 'missing' 3).  Gabow-Myers handles any digraph; so does the
    replay.  The canonical parallel search needs a DAG and says so."""
    import lichee_b200
    vaf = np.array([[0.5, 0.5, 0.5], [0.30, 0.21, 0.24], [0.21, 0.30, 0.24], [0.24, 0.24, 0.27]])
    adj = [[1, 2, 3], [2], [3], [1]]
    st, trees, ref = assert_identical_to_reference(adj, vaf, 0.1, 10)
    assert st["num_trees"] == ref["n_trees"] > 1
    with pytest.raises(lichee_b200.LicheeError) as e:
        ref_search(adj, vaf, 0.1, 10, ref_order_budget=-1)
    assert e.value.code == -3
    with pytest.raises(lichee_b200.LicheeError) as e:
        ref_search(adj, vaf, 0.1, 10, ref_order_budget=2)
    assert e.value.code == -3


def test_rerun_and_one_shot_entry_point(built):
    """A handle can be run again (the initial state is uploaded afresh); lichee_get_lineage_trees, the one call
    the JNI / Panama stub binds, returns the reference-order list."""
    import ctypes
    import lichee_b200
    net, prm = make_instance(9, 4, seed=1)
    vaf = np.ascontiguousarray(net.aaf, dtype=np.float64)
    ref = oracle.search(net.adj, vaf, prm.vaf_error_margin, top_s=50)
    s = lichee_b200.LineageSearch(net.adj, vaf, prm.vaf_error_margin, num_save=50)
    a = s.run(); ta = s.trees()
    b = s.run(); tb = s.trees()
    s.close()
    assert a.num_trees == b.num_trees == ref["n_trees"]
    assert [(t.errorScore, t.parent, t.treeNodes) for t in ta] == [(t.errorScore, t.parent, t.treeNodes) for t in tb]
    off, idx = oracle.to_csr(net.adj)
    p = lichee_b200.default_params(); p.num_save = 50; p.vaf_error_margin = prm.vaf_error_margin
    st = lichee_b200.SearchStats()
    n = vaf.shape[0]
    sc = np.zeros(50); par = np.zeros((50, n), dtype=np.int32); order = np.zeros((50, n), dtype=np.int32)
    L = lichee_b200.lib()
    L.lichee_get_lineage_trees.argtypes = [ctypes.c_int32, ctypes.c_int32] + [ctypes.c_void_p] * 8
    rc = L.lichee_get_lineage_trees(n, vaf.shape[1], off.ctypes.data, idx.ctypes.data, vaf.ctypes.data, ctypes.byref(p),
                                    ctypes.byref(st), sc.ctypes.data, par.ctypes.data, order.ctypes.data)
    assert rc == 0 and st.num_trees == ref["n_trees"] and not (st.status & 4)
    m = int(st.num_saved)
    assert list(sc[:m]) == list(ref["scores"]) and par[:m].tolist() == ref["parents"].tolist()
    assert order[:m].tolist() == ref["orders"].tolist()


def test_largest_network_state_in_global_memory(built):
    """128 nodes x 64 samples, complete DAG (8128 edges): the replay's state no longer fits shared memory next to the
    centroid matrix and is worked on in place in global memory; four 32-node groups, two samples per lane.  The
    reference's search of such a network never ends, so both sides stop at MAX_NUM_GROW_CALLS = 20000: the trees found
    until then (1295, all tied at score 0) must be the reference's, in the reference's order."""
    n, S = 128, 64
    rng = np.random.RandomState(5)
    vaf = np.full((n, S), 0.003) + 1e-5 * rng.rand(n, S)
    vaf[0, :] = 0.5
    vaf[1:8, :] = 0.02
    adj = [[j for j in range(i + 1, n)] for i in range(n)]
    for a in adj:
        rng.shuffle(a)
    st, trees, ref = assert_identical_to_reference(adj, vaf, -0.001, 50, max_trees=3000, max_grow=20000)
    assert st["num_trees"] == ref["n_trees"] > 1000 and st["status"] & 2
