#!/usr/bin/env python
"""Where the end-to-end time of a small search goes: host-side phases of LineageSearch around the
device search (median of 20), for the 120x64 network with the reference's default caps and for a
bundled dataset.  Run on a GPU box: python scripts/profile_e2e.py"""
import json, os, statistics, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lichee_b200
from lichee_b200.synth import make_instance


def phases(adj, vaf, margin, save, max_trees=100000, reps=20, budget=0):
    rows = []
    for _ in range(reps):
        t = [time.perf_counter()]
        s = lichee_b200.LineageSearch(adj, vaf, margin, num_save=save, max_num_trees=max_trees, ref_order_budget=budget); t.append(time.perf_counter())
        st = s.run(); t.append(time.perf_counter())
        tr = s.trees(); t.append(time.perf_counter())
        s.close(); t.append(time.perf_counter())
        rows.append([1e3 * (b - a) for a, b in zip(t, t[1:])] + [st.kernel_ms_total, st.num_launches])
    med = [round(statistics.median(r[i] for r in rows[min(2, len(rows) - 1):]), 3) for i in range(6)]
    out = dict(zip(["create_ms", "run_ms", "trees_ms", "close_ms", "device_ms", "launches"], med))
    out["kernel_ms"] = {k: round(st.kernel_ms[i], 3) for i, k in enumerate(st.KERNELS)}
    out["kernel_launches"] = {k: int(st.kernel_launches[i]) for i, k in enumerate(st.KERNELS)}
    return out


if __name__ == "__main__":
    net, prm = make_instance(120, 64, seed=2026)
    v = np.ascontiguousarray(net.aaf, dtype=np.float64)
    if os.environ.get("LICHEE_DEBUG"):
        phases(net.adj, v, prm.vaf_error_margin, 1000, reps=1, budget=-1)
        sys.exit(0)
    print("120x64 default caps -s 1000 (canonical):", phases(net.adj, v, prm.vaf_error_margin, 1000, budget=-1))
    net, prm = make_instance(40, 20, seed=2026)
    v = np.ascontiguousarray(net.aaf, dtype=np.float64)
    print("40x20 default caps -s 100 (canonical):", phases(net.adj, v, prm.vaf_error_margin, 100, budget=-1))
    for e in json.load(open(os.path.join(ROOT, "tests", "golden", "networks_bundled.json"))):
        if e["clustering"] == "single" and not e["complete_edges"] and e["dataset"] in ("ccRCC/RK26.txt", "hgsc/case5.txt"):
            for budget in (0, -1):
                print(e["dataset"], "reference order" if budget == 0 else "canonical",
                      phases(e["adj"], np.array(e["aaf"]), e["vaf_error_margin"], 100 if "hgsc" in e["dataset"] else 1, budget=budget))
