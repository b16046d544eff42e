"""The configurations bench.py measures, tested at the sizes it measures them.

tests/golden/canonical_digests.json holds SHA-256 digests of the ranked top-1000 list (scores, sequence numbers,
parents) that the SEQUENTIAL canonical specification in oracle/ produced for BASELINE configs[3] (40 clusters x 20
samples) and configs[4] (120 x 64), seed 2026, at the reference's default cap and at 2^22 trees.  The GPU must
reproduce every digest bit for bit -- with the lower bound on and off (LICHEE_NO_PRUNE), with and without the
single-block scheduler (LICHEE_NO_NARROW), with and without the stored-bound prefilter (LICHEE_NO_PREFILTER), with and
without the second stream that checks the next chunk ahead (LICHEE_NO_PREFETCH), and with a frontier capacity of 2^16 items -- so a bound or scheduler bug
that drops a winner cannot pass.  At 2^26 trees, where the CPU cannot follow, bound on / off must agree with each other."""
import json
import os

import numpy as np
import pytest

from lichee_b200.synth import make_instance

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "canonical_digests.json")))["cases"]
ENVS = ({}, {"LICHEE_NO_PRUNE": "1"}, {"LICHEE_NO_NARROW": "1"}, {"LICHEE_NO_PRUNE": "1", "LICHEE_NO_NARROW": "1"},
        {"LICHEE_NO_PREFILTER": "1"}, {"LICHEE_NO_PREFETCH": "1"}, {"LICHEE_NO_PREFETCH": "1", "LICHEE_NO_PRUNE": "1"})
SWITCHES = ("LICHEE_NO_NARROW", "LICHEE_NO_PRUNE", "LICHEE_NO_PREFILTER", "LICHEE_NO_PREFETCH")


def canonical_run(net, prm, cap, monkeypatch, env=None, **kw):
    import lichee_b200
    for k in SWITCHES:
        monkeypatch.delenv(k, raising=False)
    for k, v in (env or {}).items():
        monkeypatch.setenv(k, v)
    s = lichee_b200.LineageSearch(net.adj, np.array(net.aaf), prm.vaf_error_margin, num_save=1000, max_num_trees=cap,
                                  max_num_grow_calls=1 << 62, ref_order_budget=-1, **kw)
    try:
        st = s.run()
        return st, lichee_b200.top_list_sha256(s.trees())
    finally:
        s.close()
        for k in SWITCHES:
            monkeypatch.delenv(k, raising=False)


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['clusters']}x{c['samples']}-{c['max_num_trees']}")
def test_benched_shapes_match_the_pinned_digests(built, monkeypatch, case):
    net, prm = make_instance(case["clusters"], case["samples"], seed=2026)
    for env in ENVS:
        st, digest = canonical_run(net, prm, case["max_num_trees"], monkeypatch, env)
        assert st.num_trees == case["n_trees"], env
        assert digest == case["sha256"], env
        if env.get("LICHEE_NO_PRUNE"):
            assert st.num_scored == st.num_trees
    st, digest = canonical_run(net, prm, case["max_num_trees"], monkeypatch, level_capacity=1 << 16)
    assert digest == case["sha256"] and st.num_trees == case["n_trees"]


@pytest.mark.parametrize("clusters,samples", [(40, 20), (120, 64)])
def test_bound_on_and_off_agree_at_2_to_26_trees(built, monkeypatch, clusters, samples):
    net, prm = make_instance(clusters, samples, seed=2026)
    a, da = canonical_run(net, prm, 1 << 26, monkeypatch)
    b, db = canonical_run(net, prm, 1 << 26, monkeypatch, {"LICHEE_NO_PRUNE": "1"})
    c, dc = canonical_run(net, prm, 1 << 26, monkeypatch, {"LICHEE_NO_PREFETCH": "1"})
    assert a.num_trees == b.num_trees == c.num_trees == 1 << 26
    assert da == db == dc


@pytest.mark.skipif(not os.environ.get("LICHEE_SLOW"), reason="3.5*10^7 grow() calls on one warp: minutes; set LICHEE_SLOW=1")
def test_config3_identical_to_the_reference_with_a_large_budget(built):
    """BASELINE configs[3] with the reference's own caps: the reference reaches MAX_NUM_TREES after 35 363 220 grow()
    calls (tests/golden/reference_caps.json, literal port, 57 s of one core).  With a replay budget beyond that the
    library returns the reference's list: the digest of its top 100 (scores, parents, treeNodes, discovery indices)."""
    import hashlib
    import lichee_b200
    ref = json.load(open(os.path.join(HERE, "golden", "reference_caps.json")))["results"]["40x20"]
    net, prm = make_instance(40, 20, seed=2026)
    s = lichee_b200.LineageSearch(net.adj, np.array(net.aaf), prm.vaf_error_margin, num_save=100, ref_order_budget=1 << 26)
    st = s.run()
    trees = s.trees()
    s.close()
    assert not (st.status & 4) and st.status & 1
    assert (st.num_trees, st.num_grow_calls) == (ref["n_trees"], ref["n_grow"])
    h = hashlib.sha256()
    h.update(np.array([t.errorScore for t in trees], dtype=np.float64).tobytes())
    h.update(np.array([t.parent for t in trees], dtype=np.int32).tobytes())
    h.update(np.array([t.treeNodes for t in trees], dtype=np.int32).tobytes())
    h.update(np.array([t.seq for t in trees], dtype=np.int64).tobytes())
    assert h.hexdigest() == ref["top100_sha256"]


@pytest.mark.skipif(not os.environ.get("LICHEE_SLOW"), reason="10^8 grow() calls on one warp: minutes; set LICHEE_SLOW=1")
def test_config4_identical_to_the_reference_with_the_full_budget(built):
    """BASELINE configs[4] (120 x 64, -c) with the reference's own caps and a replay budget that covers them: the
    reference spends its whole MAX_NUM_GROW_CALLS = 10^8 without finding a single valid tree
    (tests/golden/reference_caps.json, literal port, 976 s of one core) -- and so does the library when it is told to
    follow the reference to the end: no tree, the grow-call cap reported, the canonical fallback not taken."""
    import lichee_b200
    ref = json.load(open(os.path.join(HERE, "golden", "reference_caps.json")))["results"]["120x64"]
    net, prm = make_instance(120, 64, seed=2026)
    s = lichee_b200.LineageSearch(net.adj, np.array(net.aaf), prm.vaf_error_margin, num_save=1000, ref_order_budget=(1 << 27))
    st = s.run()
    trees = s.trees()
    s.close()
    assert not (st.status & 4) and st.status & 2
    assert (st.num_trees, len(trees)) == (ref["n_trees"], 0) and ref["n_trees"] == 0
    assert st.num_grow_calls == ref["n_grow"] == 100000000
