"""Reference-order mode vs canonical mode vs the CPU port on the bundled networks: end-to-end wall time
through the public API (create + run + trees) and the device time of the replay."""
import json, os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lichee_b200, oracle

nets = json.load(open(os.path.join(ROOT, "tests", "golden", "networks_bundled.json")))
rows = []
for e in nets:
    if e["clustering"] != "single":
        continue
    adj, vaf = e["adj"], np.array(e["aaf"])
    rec = {"dataset": e["dataset"], "c": e["complete_edges"], "nodes": len(adj)}
    for mode, budget in (("ref", 0), ("canon", -1)):
        ts = []
        for _ in range(6):
            t0 = time.perf_counter()
            s = lichee_b200.LineageSearch(adj, vaf, e["vaf_error_margin"], num_save=100, ref_order_budget=budget)
            st = s.run(); s.trees(); s.close()
            ts.append(1e3 * (time.perf_counter() - t0))
        rec[mode + "_e2e_ms"] = round(statistics.median(ts[1:]), 3)
        rec[mode + "_dev_ms"] = round(st.kernel_ms_total, 3)
        if mode == "ref":
            rec["trees"] = int(st.num_trees); rec["grow"] = int(st.num_grow_calls); rec["launches"] = int(st.num_launches)
            rec["replay_ms"] = round(list(st.kernel_ms)[0], 3)
    t0 = time.perf_counter(); oracle.search(adj, vaf, e["vaf_error_margin"], top_s=100); rec["cpu_port_ms"] = round(1e3 * (time.perf_counter() - t0), 3)
    rows.append(rec)
    print(json.dumps(rec), flush=True)
