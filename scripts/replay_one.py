"""One reference-order search of a bundled network (default RK26 -c): the workload for profiling replay_kernel."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lichee_b200
name = sys.argv[1] if len(sys.argv) > 1 else "ccRCC/RK26.txt"
e = [x for x in json.load(open(os.path.join(ROOT, "tests", "golden", "networks_bundled.json")))
     if x["dataset"] == name and x["complete_edges"] and x["clustering"] == "single"][0]
s = lichee_b200.LineageSearch(e["adj"], np.array(e["aaf"]), e["vaf_error_margin"], num_save=100)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    t0 = time.perf_counter(); st = s.run(); dt = time.perf_counter() - t0
    print(name, "trees", st.num_trees, "grow", st.num_grow_calls, "ms", round(1e3 * dt, 2), "replay ms", round(list(st.kernel_ms)[0], 2),
          "ns/grow", round(1e6 * list(st.kernel_ms)[0] / max(st.num_grow_calls, 1), 1))
s.close()
