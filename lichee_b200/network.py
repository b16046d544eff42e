"""Pure-Python restatement of the reference code on the input side of the tree search: SSNV
loading and calling, grouping, cluster collapse and constraint-network construction (DESIGN.md row f).
Independent of the C++ host mirror (csrc/lichee_host.cpp); the tests require the two to agree bit for
bit on every bundled dataset.  Used to build fixtures and the synthetic workloads.

Reference (read-only):
  SNVDataStore.loadSNVFile / parseSNVEntry / processSNVEntry   SNVDataStore.java:385-409,522-545,663-718
  SNVDataStore small-group handling + assignAmbiguousSNVs       SNVDataStore.java:88-106,113-191
  SNVDataStore.mergeAmbiguousSNVs / filterGroups                SNVDataStore.java:197-267,347-374
  SNVGroup constructor / setSubPopulations                      SNVGroup.java:77-96,222-332
  PHYNetwork constructor / checkAndAddEdge / getAAFErrorMargin  PHYNetwork.java:94-175,185-232,237-263
  PHYNetwork.addAllHiddenEdges                                  PHYNetwork.java:302-316
  PHYNode.getAAF / getStdDev                                    PHYNode.java:161-192

Clustering itself (AAFClusterer.em -> Weka EM) is NOT restated: Weka is absent from the
reference checkout (lib/weka.jar is a missing blob) and clustering is out of scope for the
hot path. `simple_clusters` below is a documented stand-in used only to make fixtures.
"""
import math
from dataclasses import dataclass, field


@dataclass
class Params:
    max_vaf_absent: float = 0.03
    min_vaf_present: float = 0.05
    max_allowed_vaf: float = 0.6
    min_group_profile_support: int = 2
    min_snvs_per_group: int = 1
    min_robust_snvs_per_group: int = 0
    min_vaf_target_ratio: float = 0.5
    min_cluster_size: int = 2
    min_private_cluster_size: int = 1
    min_robust_cluster_support: int = 2
    max_collapse_cluster_diff: float = 0.2
    vaf_max: float = 0.5
    vaf_error_margin: float = 0.1
    all_edges: bool = False


@dataclass
class SNV:
    id: int
    line: str
    vaf: list
    profile: str = ""
    ambig: str = ""
    robust: bool = True
    desc: str = ""


@dataclass
class Cluster:
    centroid: list
    stddev: list
    members: list
    id: int = 0
    robust: bool = False


@dataclass
class Group:
    tag: str
    snvs: list
    robust: bool
    clusters: list = field(default_factory=list)

    @property
    def sample_index(self):
        return [i for i, ch in enumerate(self.tag) if ch == '1']

    @property
    def num_samples(self):
        return self.tag.count('1')


def java_string_hash(s):
    h = 0
    for ch in s:
        h = (31 * h + ord(ch)) & 0xFFFFFFFF
    return h


def java_hashmap_key_order(keys_in_insertion_order, max_size=0):
    """Iteration order of java.util.HashMap<String,?>.keySet(): buckets in index order, insertion
    order inside a bucket (bins stay linked lists for < 8 collisions; a resize preserves the
    relative order inside a bin).  The table doubles when size exceeds 0.75*capacity and never
    shrinks, so `max_size` = the largest number of keys the map ever held."""
    keys = list(keys_in_insertion_order)
    cap = 16
    while max(len(keys), max_size) > 0.75 * cap:
        cap *= 2
    def spread(h):
        return (h ^ (h >> 16)) & 0xFFFFFFFF
    buckets = {}
    for k in keys:
        buckets.setdefault(spread(java_string_hash(k)) & (cap - 1), []).append(k)
    out = []
    for b in sorted(buckets):
        out.extend(buckets[b])
    return out


def load_snv_file(path, prm: Params, normal_sample=0):
    """loadSNVFile + parseSNVEntry + processSNVEntry (format: SNV without profile)."""
    with open(path) as fh:
        lines = fh.read().split('\n')
    while lines and lines[-1] == '':
        lines.pop()
    header = lines[0].split('\t')
    sample_names = header[3:]
    S = len(sample_names)
    somatic, ambiguous = [], []
    tag2snvs = {}
    insertion = []
    for lineno, line in enumerate(lines[1:], start=1):
        parts = line.split('\t')
        assert len(parts) == 3 + S, line
        e = SNV(lineno, line, [float(x) for x in parts[3:]], desc=parts[2].strip())
        for i in range(S):
            if e.vaf[i] < prm.min_vaf_present:
                if e.vaf[i] >= prm.max_vaf_absent:
                    e.robust = False
                    e.ambig += '*'
                else:
                    e.ambig += '0'
                e.profile += '0'
            else:
                e.profile += '1'
                e.ambig += '1'
        # processSNVEntry
        if e.profile[normal_sample] == '1':
            continue
        if any(v > prm.max_allowed_vaf for v in e.vaf):
            continue
        if e.robust and '1' not in e.profile:
            continue
        somatic.append(e)
        if e.robust:
            if e.profile not in tag2snvs:
                tag2snvs[e.profile] = []
                insertion.append(e.profile)
            tag2snvs[e.profile].append(e)
        else:
            ambiguous.append(e)
    return sample_names, somatic, ambiguous, tag2snvs, insertion


def _evidence(snv, i, prm):
    return snv.vaf[i] > prm.max_vaf_absent


def _can_convert(snv, target, prm):
    for i, ch in enumerate(snv.profile):
        if ch != target[i]:
            if ch == '1':
                return False
            if not _evidence(snv, i, prm):
                return False
    return True


def build_groups(path, prm: Params, normal_sample=0):
    """SNVDataStore constructor (SNVDataStore.java:67-107). Returns (sample_names, tags in the
    order LineageEngine iterates them, tag -> SNV list, tag -> robust flag).
    The HashMap iteration order is emulated for the common no-removal case; ambiguous-SNV
    set-cover follows the reference but ties in HashMap order inside it are resolved by
    emulated key order as well."""
    names, somatic, ambiguous, tag2snvs, insertion = load_snv_file(path, prm, normal_sample)
    S = len(names)
    hwm = [len(insertion)]                       # high-water mark of tag2SNVs.size()
    def hm_order(keys):                          # iteration order of tag2SNVs.keySet()
        hwm[0] = max(hwm[0], len(keys))
        return java_hashmap_key_order(keys, hwm[0])
    # small groups -> ambiguous
    small = [t for t in hm_order(insertion)
             if len(tag2snvs[t]) < prm.min_group_profile_support]
    for t in hm_order(insertion):
        if t in small:
            ambiguous.extend(tag2snvs[t])
    for t in small:
        del tag2snvs[t]
        insertion.remove(t)
    # assignAmbiguousSNVs
    if ambiguous:
        all1, all0 = '1' * S, '0' * S
        targets = hm_order(insertion)
        if all0 not in tag2snvs:
            targets.append(all0)
        targets.append(all1)
        to_remove = []
        for snv in ambiguous:
            best, best_t = 0.0, ""
            for target in targets:
                if not _can_convert(snv, target, prm):
                    continue
                if target == all1:
                    to_remove.append(snv)
                    break
                if target == all0:
                    d = sum(0.0001 / snv.vaf[i] for i in range(S)
                            if snv.profile[i] == '1' or _evidence(snv, i, prm))
                    if d > best:
                        best, best_t = d, target
                    continue
                for ts in tag2snvs[target]:
                    d = 0.0
                    for i in range(S):
                        if ts.profile[i] == '0' and snv.profile[i] == '0' and not _evidence(snv, i, prm):
                            continue
                        mn = min(ts.vaf[i], snv.vaf[i])
                        mx = max(ts.vaf[i], snv.vaf[i])
                        if mn == 0:
                            mn = 0.0001
                        d += mn / mx
                    if d > best:
                        best, best_t = d, target
            if snv in to_remove and best_t == "":
                continue
            if best_t == all0:
                snv.profile = best_t
                continue
            if best != 0 and best_t and best >= best_t.count('1') * prm.min_vaf_target_ratio:
                snv.profile = best_t
                if snv not in to_remove:
                    to_remove.append(snv)
                tag2snvs[best_t].append(snv)
                continue
        ambiguous = [s for s in ambiguous if s not in to_remove]
        # mergeAmbiguousSNVs (greedy set cover)
        groups = {}
        ginsert = []
        adj = {}
        ainsert = []
        def extend(partial, snv, out):
            k = len(partial)
            if k == S:
                out.append(partial)
                return
            extend(partial + snv.profile[k], snv, out)
            if snv.profile[k] == '0' and _evidence(snv, k, prm):
                extend(partial + '1', snv, out)
        tl = []
        for snv in ambiguous:
            extend("", snv, tl)
        for t in tl:
            if t not in adj:
                adj[t] = []
                ainsert.append(t)
        rest = list(ambiguous)
        for snv in rest:
            for t in java_hashmap_key_order(ainsert):
                if _can_convert(snv, t, prm):
                    adj[t].append(snv)
        while rest:
            mx, mset = 0, ""
            for t in java_hashmap_key_order([x for x in ainsert if x in adj], len(ainsert)):
                if len(adj[t]) > mx:
                    mx, mset = len(adj[t]), t
            if mx <= 1:
                break
            groups[mset] = adj[mset]
            ginsert.append(mset)
            for entry in list(adj[mset]):
                entry.profile = mset
                for t in adj:
                    if t == mset:
                        continue
                    if entry in adj[t]:
                        adj[t].remove(entry)
                if entry in rest:
                    rest.remove(entry)
            del adj[mset]
        for snv in rest:
            tag = ""
            for i in range(S):
                if snv.profile[i] == '0' and _evidence(snv, i, prm):
                    d0 = snv.vaf[i] - prm.max_vaf_absent
                    d1 = prm.min_vaf_present - snv.vaf[i]
                    tag += '1' if d1 < d0 else snv.profile[i]
                else:
                    tag += snv.profile[i]
            if tag not in groups:
                groups[tag] = []
                ginsert.append(tag)
            groups[tag].append(snv)
        for t in java_hashmap_key_order(ginsert):
            if t == all0:
                continue
            if t not in tag2snvs:
                tag2snvs[t] = groups[t]
                insertion.append(t)
        if all0 in tag2snvs:
            del tag2snvs[all0]
            insertion.remove(all0)
    # filterGroups
    for t in list(insertion):
        if '1' not in t or len(tag2snvs[t]) < prm.min_snvs_per_group:
            del tag2snvs[t]
            insertion.remove(t)
    order = hm_order(insertion)
    robust = {}
    for t in order:
        nr = sum(1 for e in tag2snvs[t] if e.robust)
        robust[t] = nr >= prm.min_group_profile_support
    return names, order, tag2snvs, robust


def group_matrix(g: Group):
    idx = g.sample_index
    return [[snv.vaf[i] for i in idx] for snv in g.snvs]


def recompute(c: Cluster, data, nfeat):
    m = len(c.members)
    cen = [0.0] * nfeat
    for i in c.members:
        for j in range(nfeat):
            cen[j] += data[i][j]
    cen = [x / m for x in cen]
    sd = [0.0] * nfeat
    for i in c.members:
        for j in range(nfeat):
            sd[j] += (data[i][j] - cen[j]) ** 2
    sd = [math.sqrt(x / m) for x in sd]
    c.centroid, c.stddev = cen, sd


def simple_clusters(g: Group, gap=0.08):
    """STAND-IN for AAFClusterer.em (Weka): deterministic 1-D chaining on the mean VAF,
    splitting where consecutive sorted means differ by more than `gap`."""
    data = group_matrix(g)
    nf = g.num_samples
    means = sorted(range(len(data)), key=lambda i: (sum(data[i]) / nf, i))
    parts, cur = [], [means[0]]
    for a, b in zip(means, means[1:]):
        if sum(data[b]) / nf - sum(data[a]) / nf > gap:
            parts.append(cur)
            cur = []
        cur.append(b)
    parts.append(cur)
    out = []
    for k, mem in enumerate(parts):
        c = Cluster([], [], sorted(mem), k)
        recompute(c, data, nf)
        out.append(c)
    return out


def set_sub_populations(g: Group, clusters, prm: Params):
    """SNVGroup.setSubPopulations (SNVGroup.java:222-332): size filter, greedy collapse of
    close centroids (AVG_PER_SAMPLE distance), robustness flag."""
    data = group_matrix(g)
    nf = g.num_samples
    filt = [c for c in clusters
            if len(c.members) >= prm.min_cluster_size
            or (nf == 1 and len(c.members) >= prm.min_private_cluster_size)]
    if not filt:
        g.clusters = []
        return
    def dist(a, b):
        return sum(abs(x - y) for x, y in zip(a.centroid, b.centroid)) / len(a.centroid)
    q = []
    def qinsert(pd):
        k = 0
        while k < len(q) and not (q[k][2] > pd[2]):
            k += 1
        q.insert(k, pd)
    for i in range(len(filt)):
        for j in range(i + 1, len(filt)):
            qinsert((filt[i].id, filt[j].id, dist(filt[i], filt[j])))
    byid = {c.id: c for c in clusters}
    while q and q[0][2] < prm.max_collapse_cluster_diff:
        id1, id2, _ = q.pop(0)
        c1, c2 = byid[id1], byid[id2]
        c1.members.extend(c2.members)
        recompute(c1, data, nf)
        filt.remove(c2)
        q[:] = [x for x in q if x[0] not in (id1, id2) and x[1] not in (id1, id2)]
        for c in filt:
            if c.id in (id1, id2):
                continue
            qinsert((c1.id, c.id, dist(c1, c)))
    for c in filt:
        nr = 0
        for m in c.members:
            if g.snvs[m].robust:
                nr += 1
                if nr >= prm.min_robust_cluster_support:
                    c.robust = True
                    break
    g.clusters = filt


class Network:
    """PHYNetwork constructor. Node 0 = root; nodes numbered in creation order.
    aaf[node][sample], sd[node][sample], size[node], level[node]; adj = child lists in
    reference insertion order."""

    def __init__(self, groups, S, prm: Params):
        self.S = S
        self.prm = prm
        self.aaf = [[prm.vaf_max] * S]
        self.sd = [[0.0] * S]
        self.size = [0]
        self.level = [S + 1]
        self.is_root = [True]
        self.info = [None]
        self.adj = [[]]
        self.levels = {S + 1: [0]}
        self.num_edges = 0
        for g in groups:
            ids = []
            idx = g.sample_index
            for ci, c in enumerate(g.clusters):
                row = [0.0] * S
                sdr = [0.0] * S
                for k, s in enumerate(idx):
                    row[s] = c.centroid[k]
                    sdr[s] = c.stddev[k] if c.stddev else 0.0
                nid = len(self.aaf)
                self.aaf.append(row); self.sd.append(sdr); self.size.append(len(c.members))
                self.level.append(g.num_samples); self.is_root.append(False)
                self.info.append((g.tag, ci)); self.adj.append([])
                self.levels.setdefault(g.num_samples, []).append(nid)
                ids.append(nid)
            for i in range(len(ids)):
                for j in range(i + 1, len(ids)):
                    self.check_and_add(ids[i], ids[j])
        for i in range(S + 1, 0, -1):
            fr = self.levels.get(i)
            if fr is None:
                continue
            j = i - 1
            to = self.levels.get(j)
            while to is None and j > 0:
                j -= 1
                to = self.levels.get(j)
            if to is None:
                continue
            for n1 in fr:
                for n2 in to:
                    self.check_and_add(n1, n2)
        if prm.all_edges:
            for i in range(S + 1, 0, -1):
                fr = self.levels.get(i)
                if fr is None:
                    continue
                for j in range(i - 1, 0, -1):
                    to = self.levels.get(j)
                    if to is None:
                        continue
                    for n1 in fr:
                        for n2 in to:
                            self.check_and_add(n1, n2)
        n = len(self.aaf)
        mask = [0] * n
        # HashMap<PHYNode,..> iteration: order irrelevant for the mask
        for a in range(n):
            for b in self.adj[a]:
                mask[b] = 1
        for i in range(1, n):
            if mask[i] == 0:
                found = False
                for j in range(self.level[i] + 2, S + 2):
                    fr = self.levels.get(j)
                    if fr is None:
                        continue
                    for n2 in fr:
                        if self.check_and_add(n2, i) == 0:
                            found = True
                            break
                    if found:
                        break
                if not found:
                    self.add_edge(0, i)
        self.n = n

    def margin(self, a, b, i):
        p = self.prm
        if self.is_root[a]:
            pe = p.vaf_error_margin
        else:
            pe = 1.96 * self.sd[a][i] / math.sqrt(float(self.size[a]))
        if self.is_root[b]:
            ce = p.vaf_error_margin
        else:
            ce = 1.96 * self.sd[b][i] / math.sqrt(float(self.size[b]))
        se = pe + ce
        return se if se > p.vaf_error_margin else p.vaf_error_margin

    def add_edge(self, a, b):
        if b not in self.adj[a]:
            self.adj[a].append(b)
            self.num_edges += 1

    def check_and_add(self, n1, n2):
        S = self.S
        A, B = self.aaf[n1], self.aaf[n2]
        c12 = c21 = 0
        e12 = e21 = 0.0
        for i in range(S):
            if A[i] == 0 and B[i] != 0:
                break
            c12 += 1 if A[i] >= (B[i] - self.margin(n1, n2, i)) else 0
            if A[i] < B[i]:
                e12 += B[i] - A[i]
        for i in range(S):
            if B[i] == 0 and A[i] != 0:
                break
            c21 += 1 if B[i] >= (A[i] - self.margin(n2, n1, i)) else 0
            if B[i] < A[i]:
                e21 += A[i] - B[i]
        if c12 == S:
            if c21 == S:
                if e12 < e21:
                    self.add_edge(n1, n2); return 0
                self.add_edge(n2, n1); return 1
            self.add_edge(n1, n2); return 0
        elif c21 == S:
            self.add_edge(n2, n1); return 1
        return -1


def build_network_from_file(path, prm: Params, normal_sample=0, gap=0.08):
    names, order, tag2snvs, robust = build_groups(path, prm, normal_sample)
    groups = []
    for t in order:
        g = Group(t, tag2snvs[t], robust[t])
        set_sub_populations(g, simple_clusters(g, gap), prm)
        groups.append(g)
    return Network(groups, len(names), prm), groups, names
