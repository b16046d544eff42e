"""Generates tests/golden/canonical_digests.json: the sequential canonical specification (oracle/,
lichee_canonical_search) on the two shapes bench.py measures -- BASELINE configs[3] (40 clusters x 20 samples) and
configs[4] (120 x 64), seed 2026, -s 1000 -- at 2^22 trees and at the reference's default cap, as SHA-256 digests of
the ranked list (score bits, seq, parents).  The GPU must reproduce them bit for bit with the lower bound on and off,
with and without the single-block scheduler, and with a small frontier capacity (tests/test_gpu_bench_config.py);
bench.py prints the same digest of every step.  Several minutes of one core."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle
from lichee_b200.synth import make_instance


def ranked_digest(scores, seqs, parents):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(scores, dtype=np.float64).tobytes())
    h.update(np.ascontiguousarray(seqs, dtype=np.int64).tobytes())
    h.update(np.ascontiguousarray(parents, dtype=np.int32).tobytes())
    return h.hexdigest()


if __name__ == "__main__":
    out = {"seed": 2026, "num_save": 1000, "digest": "sha256(scores f64 | seq i64 | parents i32), ranked order", "cases": []}
    for (c, s) in ((40, 20), (120, 64)):
        net, prm = make_instance(c, s, seed=2026)
        for cap in (100000, 1 << 22):
            t0 = time.time()
            r = oracle.canonical_search(net.adj, np.array(net.aaf), prm.vaf_error_margin, top_s=1000, max_trees=cap)
            rec = {"clusters": c, "samples": s, "max_num_trees": cap, "n_trees": r["n_trees"], "n_checks": r["n_checks"],
                   "n_grow": r["n_grow"], "best_score": float(r["scores"][0]), "worst_kept_score": float(r["scores"][-1]),
                   "sha256": ranked_digest(r["scores"], r["seq"], r["parents"]), "cpu_seconds": round(time.time() - t0, 2)}
            out["cases"].append(rec)
            print(rec, flush=True)
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "canonical_digests.json"), "w"), indent=1)
