"""Generates tests/golden/*.json.  Run HERE (the container that has /root/reference); the tests
and the GPU box only read the committed JSON.

  networks_bundled.json : constraint networks of the bundled ccRCC / HGSC datasets with the
      README / data/README settings (BASELINE.json configs[0..2]), built with
      lichee_b200.network (restatement of SNVDataStore / SNVGroup / PHYNetwork constructor).
      Clustering is a stand-in for Weka EM: `single` = one cluster per SSNV group (what EM +
      maxClusterDist collapse yields on these data according to the published figures),
      `split` = 1-D chaining with gap 0.08 (exercises networks with several clusters per group).
      Each entry also stores what the literal reference search (oracle mode 0) returns on it, so
      the fixture doubles as a regression pin for the oracle itself.
  published_trees.json : the two top trees the reference publishes as figures
      (img_demo/RK26.txt.dot.png, img_demo/RMH008.txt.dot.png), transcribed by hand: nodes are
      identified by the set of samples their SSNVs are present in, and labelled with the cluster
      size printed in the figure.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import oracle
from lichee_b200.network import Params, build_network_from_file

D = "/root/reference/LICHeE/data/"
CFG = {
    "ccRCC/RK26.txt": (0.005, 0.005, {}), "ccRCC/RMH008.txt": (0.005, 0.005, {"min_private_cluster_size": 2}),
    "ccRCC/EV003.txt": (0.005, 0.005, {}), "ccRCC/EV005.txt": (0.005, 0.005, {}), "ccRCC/EV006.txt": (0.005, 0.005, {}),
    "ccRCC/EV007.txt": (0.005, 0.005, {}), "ccRCC/RMH002.txt": (0.005, 0.005, {}), "ccRCC/RMH004.txt": (0.01, 0.01, {}),
    "hgsc/case1.txt": (0.005, 0.01, {}), "hgsc/case2.txt": (0.005, 0.01, {}), "hgsc/case3.txt": (0.005, 0.01, {}),
    "hgsc/case4.txt": (0.005, 0.01, {"min_cluster_size": 3}),
    "hgsc/case5.txt": (0.01, 0.04, {"min_cluster_size": 3, "max_collapse_cluster_diff": 0.1}),
    "hgsc/case6.txt": (0.005, 0.01, {"max_collapse_cluster_diff": 0.1}),
}

out = []
for f, (ab, pr, kw) in CFG.items():
    for mode, gap in (("single", 1e9), ("split", 0.08)):
        for complete in (False, True):
            prm = Params(max_vaf_absent=ab, min_vaf_present=pr, all_edges=complete, **kw)
            net, groups, names = build_network_from_file(D + f, prm, gap=gap)
            r = oracle.search(net.adj, net.aaf, prm.vaf_error_margin, top_s=100, want_all=100000)
            out.append({
                "dataset": f, "clustering": mode, "complete_edges": complete, "samples": names,
                "vaf_error_margin": prm.vaf_error_margin, "adj": net.adj, "aaf": [[float(x) for x in row] for row in net.aaf],
                "tags": [None] + [net.info[i][0] for i in range(1, net.n)], "sizes": net.size,
                "reference_search": {"n_trees": r["n_trees"], "n_grow": r["n_grow"], "n_checks": r["n_checks"],
                                     "scores": [float(x) for x in r["scores"]],
                                     "parents": [[int(x) for x in p] for p in r["parents"]],
                                     "orders": [[int(x) for x in p] for p in r["orders"]],
                                     "all_scores_sorted": sorted(float(x) for x in r["all_scores"])[:300]},
            })
json.dump(out, open(os.path.join(os.path.dirname(__file__), "networks_bundled.json"), "w"))
print(len(out), "networks;", sum(e["reference_search"]["n_trees"] for e in out), "valid trees in total")

# Published figures, transcribed: sample order is the dataset's header order.
RK26 = ["Normal", "R2", "R1", "R3", "R4", "R11", "R10", "R9", "R5", "R6", "R7", "R8"]
RMH008 = ["Normal", "R1", "R2", "R3", "R6", "R4", "R5", "R7", "R8"]
def tag(order, present):
    return "".join("1" if s in present else "0" for s in order)
pub = [
    {"dataset": "ccRCC/RK26.txt", "figure": "img_demo/RK26.txt.dot.png", "clustering": "single", "complete_edges": False,
     "nodes": {tag(RK26, RK26[1:]): 6, tag(RK26, ["R2", "R1", "R3", "R4"]): 21, tag(RK26, ["R2"]): 4,
               tag(RK26, ["R1", "R3", "R4"]): 2, tag(RK26, ["R10", "R9", "R5"]): 4, tag(RK26, ["R10", "R9"]): 3,
               tag(RK26, ["R9"]): 1, tag(RK26, ["R10"]): 1, tag(RK26, ["R5"]): 2, tag(RK26, ["R5", "R6", "R7", "R8"]): 3,
               tag(RK26, ["R8"]): 6, tag(RK26, ["R11"]): 10},
     "edges": [["GL", tag(RK26, RK26[1:])],
               [tag(RK26, RK26[1:]), tag(RK26, ["R2", "R1", "R3", "R4"])], [tag(RK26, RK26[1:]), tag(RK26, ["R10", "R9", "R5"])],
               [tag(RK26, RK26[1:]), tag(RK26, ["R5", "R6", "R7", "R8"])], [tag(RK26, RK26[1:]), tag(RK26, ["R11"])],
               [tag(RK26, ["R2", "R1", "R3", "R4"]), tag(RK26, ["R2"])], [tag(RK26, ["R2", "R1", "R3", "R4"]), tag(RK26, ["R1", "R3", "R4"])],
               [tag(RK26, ["R10", "R9", "R5"]), tag(RK26, ["R10", "R9"])], [tag(RK26, ["R10", "R9", "R5"]), tag(RK26, ["R5"])],
               [tag(RK26, ["R10", "R9"]), tag(RK26, ["R9"])], [tag(RK26, ["R10", "R9"]), tag(RK26, ["R10"])],
               [tag(RK26, ["R5", "R6", "R7", "R8"]), tag(RK26, ["R8"])]]},
    {"dataset": "ccRCC/RMH008.txt", "figure": "img_demo/RMH008.txt.dot.png", "clustering": "single", "complete_edges": False,
     "nodes": {tag(RMH008, RMH008[1:]): 22, tag(RMH008, ["R1", "R2", "R3", "R6", "R4"]): 10, tag(RMH008, ["R6", "R4", "R5", "R7", "R8"]): 8,
               tag(RMH008, ["R8"]): 6, tag(RMH008, ["R5", "R7", "R4", "R6"]): 5, tag(RMH008, ["R4"]): 7, tag(RMH008, ["R6", "R3"]): 3,
               tag(RMH008, ["R6"]): 4, tag(RMH008, ["R2"]): 9, tag(RMH008, ["R1"]): 2},
     "edges": [["GL", tag(RMH008, RMH008[1:])],
               [tag(RMH008, RMH008[1:]), tag(RMH008, ["R6", "R4", "R5", "R7", "R8"])], [tag(RMH008, RMH008[1:]), tag(RMH008, ["R1", "R2", "R3", "R6", "R4"])],
               [tag(RMH008, ["R6", "R4", "R5", "R7", "R8"]), tag(RMH008, ["R8"])], [tag(RMH008, ["R6", "R4", "R5", "R7", "R8"]), tag(RMH008, ["R5", "R7", "R4", "R6"])],
               [tag(RMH008, ["R1", "R2", "R3", "R6", "R4"]), tag(RMH008, ["R4"])], [tag(RMH008, ["R1", "R2", "R3", "R6", "R4"]), tag(RMH008, ["R6", "R3"])],
               [tag(RMH008, ["R1", "R2", "R3", "R6", "R4"]), tag(RMH008, ["R2"])], [tag(RMH008, ["R1", "R2", "R3", "R6", "R4"]), tag(RMH008, ["R1"])],
               [tag(RMH008, ["R6", "R3"]), tag(RMH008, ["R6"])]]},
]
json.dump(pub, open(os.path.join(os.path.dirname(__file__), "published_trees.json"), "w"), indent=1)
# report how the transcription compares with what the restated pipeline builds
nets = {(e["dataset"], e["clustering"], e["complete_edges"]): e for e in out}
for p in pub:
    e = nets[(p["dataset"], p["clustering"], p["complete_edges"])]
    got = {e["tags"][i]: e["sizes"][i] for i in range(1, len(e["tags"]))}
    par = e["reference_search"]["parents"][0]
    edges = sorted(["GL" if par[v] == 0 else e["tags"][par[v]], e["tags"][v]] for v in range(1, len(par)))
    print(p["dataset"], "sizes equal:", got == p["nodes"], "edges equal:", edges == sorted(p["edges"]),
          {k: (got.get(k), p["nodes"][k]) for k in p["nodes"] if got.get(k) != p["nodes"][k]})
