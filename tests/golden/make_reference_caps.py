"""Generates tests/golden/reference_caps.json: the LITERAL port of the reference search (oracle/, mode 0) run with the
reference's own caps (MAX_NUM_TREES 100000, MAX_NUM_GROW_CALLS 10^8) on the two synthetic shapes BASELINE.json names
(configs[3] 40 clusters x 20 samples, configs[4] 120 x 64; seed 2026, the bench.py network).  Takes ~17 minutes of one
core: 40x20 reaches the tree cap after 3.5*10^7 grow() calls, 120x64 runs into the grow-call cap without a single
valid tree."""
import sys, time, json
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
import numpy as np
import oracle
from lichee_b200.synth import make_instance
out = {}
for (c, s) in ((40, 20), (120, 64)):
    net, prm = make_instance(c, s, seed=2026)
    t0 = time.time()
    r = oracle.search(net.adj, np.array(net.aaf), prm.vaf_error_margin, top_s=1, max_trees=100000, max_grow=100000000)
    dt = time.time() - t0
    out[f"{c}x{s}"] = dict(n_trees=r["n_trees"], n_grow=r["n_grow"], n_checks=r["n_checks"], stopped=r["stopped"], seconds=round(dt, 1))
    print(out, flush=True)
json.dump(out, open(__import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "reference_caps.json"), "w"), indent=1)

# top-100 digest of the 40x20 run (second pass with -s 100; same search)
import hashlib
net, prm = make_instance(40, 20, seed=2026)
r = oracle.search(net.adj, np.array(net.aaf), prm.vaf_error_margin, top_s=100, max_trees=100000, max_grow=100000000)
h = hashlib.sha256()
h.update(np.ascontiguousarray(r["scores"], dtype=np.float64).tobytes())
h.update(np.ascontiguousarray(r["parents"], dtype=np.int32).tobytes())
h.update(np.ascontiguousarray(r["orders"], dtype=np.int32).tobytes())
h.update(np.ascontiguousarray(r["dfs_index"], dtype=np.int64).tobytes())
print("40x20 top100_sha256", h.hexdigest(), "best", float(r["scores"][0]), "worst kept", float(r["scores"][-1]))
