"""Synthetic clonal-evolution instances of the shapes BASELINE.json names
(workload generator for tests and bench.py): a random true lineage tree over `n_clusters` clusters and
`n_samples` samples, cluster centroid VAFs obeying the sum rule up to noise, turned into a
constraint network with the reference's own edge rule (lichee_b200.network.Network,
PHYNetwork.java:94-175) under -c (ALL_EDGES) and -minClusterSize 1."""
import random
from .network import Cluster, Group, Network, Params


def make_instance(n_clusters, n_samples, seed, noise=0.01, margin=0.1, complete=True,
                  keep_prob=0.75, min_frac=0.15):
    rng = random.Random(seed)
    S = n_samples
    # true tree: node 0 = germline with VAF 0.5 everywhere
    parent = [-1]
    present = [set(range(S))]
    vaf = [[0.5] * S]
    kids = [[]]
    room = [[0.5] * S]                    # remaining VAF mass a node can hand to children
    while len(parent) < n_clusters + 1:
        p = rng.randrange(len(parent))
        cand = [s for s in present[p] if room[p][s] > 0.02]
        if not cand:
            continue
        if p == 0 and len(kids[0]) == 0:
            pres = set(cand)              # first clone is clonal in every sample
        else:
            pres = {s for s in cand if rng.random() < keep_prob}
            if not pres:
                pres = {rng.choice(cand)}
        row = [0.0] * S
        for s in pres:
            take = room[p][s] * rng.uniform(min_frac, 0.9)
            row[s] = take
            room[p][s] -= take
        parent.append(p); present.append(pres); vaf.append(row); kids.append([]); kids[p].append(len(parent) - 1)
        room.append(list(row))
    # observed centroids = truth + noise (clipped), one cluster per node; group by profile
    groups = {}
    order = []
    for v in range(1, n_clusters + 1):
        tag = ''.join('1' if s in present[v] else '0' for s in range(S))
        cen = [min(0.5, max(1e-4, vaf[v][s] + rng.gauss(0, noise))) for s in sorted(present[v])]
        sd = [abs(rng.gauss(0, noise)) for _ in cen]
        size = rng.randint(1, 30)
        if tag not in groups:
            groups[tag] = Group(tag, [], True)
            order.append(tag)
        g = groups[tag]
        g.clusters.append(Cluster(cen, sd, list(range(size)), len(g.clusters), True))
    prm = Params(vaf_error_margin=margin, all_edges=complete, min_cluster_size=1)
    rng.shuffle(order)
    net = Network([groups[t] for t in order], S, prm)
    return net, prm
