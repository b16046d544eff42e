"""One canonical search of the headline network (120 x 64, 2^30 trees): the workload for ncu captures."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lichee_b200
from lichee_b200.synth import make_instance
net, prm = make_instance(120, 64, seed=2026)
s = lichee_b200.LineageSearch(net.adj, np.ascontiguousarray(net.aaf, dtype=np.float64), prm.vaf_error_margin, num_save=1000,
                              max_num_trees=1 << 30, max_num_grow_calls=1 << 62, ref_order_budget=-1)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    t0 = time.perf_counter(); st = s.run(); dt = time.perf_counter() - t0
    print("trees", st.num_trees, "ms", round(1e3 * dt, 2), {k: round(st.kernel_ms[i], 2) for i, k in enumerate(st.KERNELS)}, "launches", st.num_launches)
s.close()
