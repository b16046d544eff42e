// lichee_host.cpp -- see lichee_host.hpp.  Reference line numbers refer to LICHeE/src/lineage.
#include "lichee_host.hpp"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <set>
#include <sstream>
#include <stdexcept>

namespace lichee {

// ---------------------------------------------------------------------------------------------
// Java runtime behaviours the output depends on
// ---------------------------------------------------------------------------------------------
int32_t javaStringHash(const std::string &s) {
    uint32_t h = 0;
    for (unsigned char c : s) h = 31u * h + c;
    return (int32_t)h;
}

std::vector<size_t> javaHashOrder(const std::vector<int32_t> &hashes, size_t max_size) {
    size_t cap = 16;
    while ((double)std::max(hashes.size(), max_size) > 0.75 * (double)cap) cap *= 2;
    std::vector<size_t> idx(hashes.size());
    for (size_t i = 0; i < idx.size(); i++) idx[i] = i;
    auto bucket = [&](size_t i) { uint32_t h = (uint32_t)hashes[i]; return (size_t)((h ^ (h >> 16)) & (cap - 1)); };
    std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return bucket(a) < bucket(b); });
    return idx;
}

std::string javaDoubleToString(double d) {
    if (std::isnan(d)) return "NaN";
    if (std::isinf(d)) return d > 0 ? "Infinity" : "-Infinity";
    if (d == 0) return std::signbit(d) ? "-0.0" : "0.0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), std::fabs(d), std::chars_format::scientific);
    std::string sci(buf, r.ptr);                     // d[.ddd]e[+-]XX, shortest round-trip digits
    size_t epos = sci.find('e');
    std::string digits;
    for (char c : sci.substr(0, epos)) if (c != '.') digits += c;
    int exp10 = std::atoi(sci.c_str() + epos + 1);
    std::string out = d < 0 ? "-" : "";
    if (exp10 >= -3 && exp10 < 7) {
        if (exp10 >= 0) {
            std::string ip = digits.substr(0, std::min<size_t>(digits.size(), exp10 + 1));
            while ((int)ip.size() < exp10 + 1) ip += '0';
            std::string fp = (int)digits.size() > exp10 + 1 ? digits.substr(exp10 + 1) : "0";
            out += ip + "." + fp;
        } else {
            out += "0." + std::string(-exp10 - 1, '0') + digits;
        }
    } else {
        out += digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0") + "E" + std::to_string(exp10);
    }
    return out;
}

std::string decimalFormat(double d, int max_frac) {
    // DecimalFormat("#.##"): HALF_EVEN, no leading zero for |d| < 1, "0" when nothing is left
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.*f", max_frac, std::fabs(d));
    std::string s(buf);
    while (!s.empty() && s.back() == '0') s.pop_back();
    if (!s.empty() && s.back() == '.') s.pop_back();
    if (s.size() > 1 && s[0] == '0' && s[1] == '.') s.erase(0, 1);
    if (s.empty() || s == "0") return "0";
    return (d < 0 ? "-" : "") + s;
}

static std::vector<std::string> split(const std::string &s, char sep) {
    // String.split: trailing empty strings are removed
    std::vector<std::string> out;
    std::string cur;
    for (char c : s) {
        if (c == sep) { out.push_back(cur); cur.clear(); } else cur += c;
    }
    out.push_back(cur);
    while (!out.empty() && out.back().empty()) out.pop_back();
    return out;
}
static std::string trim(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && (unsigned char)s[a] <= ' ') a++;
    while (b > a && (unsigned char)s[b - 1] <= ' ') b--;
    return s.substr(a, b - a);
}
[[noreturn]] static void formatError(const std::string &desc, const std::string &line) {
    std::string m = "[Wrong input file format] " + desc;
    if (!line.empty()) m += "\nLine: " + line;
    throw std::runtime_error(m);
}

// ---------------------------------------------------------------------------------------------
// Cluster / SNVGroup
// ---------------------------------------------------------------------------------------------
void Cluster::recomputeCentroidAndStdDev(const std::vector<std::vector<double>> &data, int nf) {   // AAFClusterer.java:426-448
    std::vector<double> c(nf, 0.0), sd(nf, 0.0);
    for (int m : members) for (int j = 0; j < nf; j++) c[j] += data[m][j];
    for (int j = 0; j < nf; j++) c[j] = c[j] / (double)members.size();
    for (int m : members) for (int j = 0; j < nf; j++) sd[j] += std::pow(data[m][j] - c[j], 2);
    for (int j = 0; j < nf; j++) sd[j] = std::sqrt(sd[j] / (double)members.size());
    centroid = c;
    stdDev = sd;
}

SNVGroup::SNVGroup(const std::string &t, const std::vector<SNVEntry *> &s, bool robust)      // SNVGroup.java:77-96
    : tag(t), snvs(s), isRobust(robust) {
    for (size_t i = 0; i < tag.size(); i++) if (tag[i] == '1') sampleIndex.push_back((int)i);
    for (SNVEntry *e : snvs) {
        std::vector<double> row;
        for (int si : sampleIndex) row.push_back(e->VAF[si]);
        alleleFreqBySample.push_back(row);
    }
}
int SNVGroup::getSampleIndex(int sampleId) const {
    for (size_t i = 0; i < sampleIndex.size(); i++) if (sampleIndex[i] == sampleId) return (int)i;
    return -1;
}
void SNVGroup::removeCluster(const Cluster *c) {                                                // SNVGroup.java:343-356
    std::vector<std::shared_ptr<Cluster>> keep;
    for (auto &p : subPopulations) if (p.get() != c) keep.push_back(p);
    subPopulations = keep;
}
void SNVGroup::setSubPopulations(std::vector<std::shared_ptr<Cluster>> clusters, const Parameters &p) {  // :222-332
    const int nf = getNumSamples();
    std::vector<std::shared_ptr<Cluster>> filt;
    for (auto &c : clusters)
        if ((int)c->members.size() >= p.MIN_CLUSTER_SIZE || (nf == 1 && (int)c->members.size() >= p.MIN_PRIVATE_CLUSTER_SIZE))
            filt.push_back(c);
    if (filt.empty()) { subPopulations.clear(); return; }
    struct PD { int a, b; double d; };
    std::vector<PD> q;
    auto dist = [&](const Cluster &x, const Cluster &y) {
        double s = 0;
        for (size_t i = 0; i < x.centroid.size(); i++) s += std::fabs(x.centroid[i] - y.centroid[i]);
        return s / (double)x.centroid.size();
    };
    auto qinsert = [&](PD pd) {
        size_t k = 0;
        while (k < q.size() && !(q[k].d > pd.d)) k++;
        q.insert(q.begin() + k, pd);
    };
    for (size_t i = 0; i < filt.size(); i++)
        for (size_t j = i + 1; j < filt.size(); j++) qinsert({filt[i]->id, filt[j]->id, dist(*filt[i], *filt[j])});
    auto byId = [&](int id) -> std::shared_ptr<Cluster> { return clusters[id]; };   // clusters[pd.clusterId] in the reference
    while (!q.empty() && q[0].d < p.MAX_COLLAPSE_CLUSTER_DIFF) {
        PD pd = q[0];
        q.erase(q.begin());
        auto c1 = byId(pd.a), c2 = byId(pd.b);
        for (int m : c2->members) c1->members.push_back(m);
        c1->recomputeCentroidAndStdDev(alleleFreqBySample, nf);
        filt.erase(std::remove(filt.begin(), filt.end(), c2), filt.end());
        q.erase(std::remove_if(q.begin(), q.end(), [&](const PD &x) {
                    return x.a == c1->id || x.b == c1->id || x.a == c2->id || x.b == c2->id; }), q.end());
        for (auto &c : filt) {
            if (c->id == c1->id || c->id == c2->id) continue;
            qinsert({c1->id, c->id, dist(*c1, *c)});
        }
    }
    for (auto &c : filt) {
        int nr = 0;
        for (int m : c->members)
            if (snvs[m]->isRobust && ++nr >= p.MIN_ROBUST_CLUSTER_SUPPORT) { c->robust = true; break; }
    }
    subPopulations = filt;
}

// ---------------------------------------------------------------------------------------------
// SNVDataStore
// ---------------------------------------------------------------------------------------------
void SNVDataStore::put(const std::string &tag) {
    if (!tag2SNVs.count(tag)) { tag2SNVs[tag]; insertion.push_back(tag); hwm = std::max(hwm, insertion.size()); }
}
void SNVDataStore::erase(const std::string &tag) {
    tag2SNVs.erase(tag);
    insertion.erase(std::remove(insertion.begin(), insertion.end(), tag), insertion.end());
}
std::vector<std::string> SNVDataStore::groupTags() const {
    std::vector<int32_t> h;
    for (auto &t : insertion) h.push_back(javaStringHash(t));
    std::vector<std::string> out;
    for (size_t i : javaHashOrder(h, hwm)) out.push_back(insertion[i]);
    return out;
}
bool SNVDataStore::isRobustGroup(const std::string &tag) const {                               // :460-471
    int n = 0;
    for (SNVEntry *e : tag2SNVs.at(tag)) if (e->isRobust && ++n >= prm.MIN_GROUP_PROFILE_SUPPORT) return true;
    return false;
}
void SNVDataStore::parseHeader(const std::string &line) {                                      // :638-647
    const int nreq = prm.INPUT_FORMAT == Parameters::SNV ? 3 : 4;
    auto parts = split(line, '\t');
    if ((int)parts.size() < nreq + 2)
        formatError("The header must have " + std::to_string(nreq) + " required fields and at least 2 samples (separated by tabs)", line);
    sampleNames.assign(parts.begin() + nreq, parts.end());
}
std::unique_ptr<SNVEntry> SNVDataStore::parseSNVEntry(const std::string &line, int lineId) {   // :663-718
    auto e = std::make_unique<SNVEntry>();
    e->id = lineId;
    e->line = line;
    auto parts = split(line, '\t');
    const int nf = prm.INPUT_FORMAT == Parameters::SNV ? 3 : 4;
    const int S = getNumSamples();
    if ((int)parts.size() != nf + S)
        formatError("Expecting " + std::to_string(nf) + " fields and " + std::to_string(S) + " samples based on the header", line);
    std::string chr = trim(parts[0]);
    std::transform(chr.begin(), chr.end(), chr.begin(), ::tolower);
    if (chr.find("chr") != std::string::npos) chr = chr.substr(3);
    if (chr == "x") e->chr = 23;
    else if (chr == "y") e->chr = 24;
    else {
        char *end = nullptr;
        long v = std::strtol(chr.c_str(), &end, 10);
        if (chr.empty() || *end) formatError("Chromosome " + parts[0], line);
        e->chr = (int)v;
    }
    {
        std::string p = trim(parts[1]);
        char *end = nullptr;
        long v = std::strtol(p.c_str(), &end, 10);
        if (p.empty() || *end) formatError("Position " + p, line);
        e->position = (int)v;
    }
    e->description = trim(parts[2]);
    if (prm.INPUT_FORMAT == Parameters::SNV_WITH_PROFILE) {
        e->presenceProfile = trim(parts[3]);
        if ((int)e->presenceProfile.size() != S)
            formatError("Presence profile " + e->presenceProfile + " length does not match the number of input samples", line);
    }
    e->VAF.resize(S);
    e->isRobust = true;
    for (int i = 0; i < S; i++) {
        const std::string &f = parts[nf + i];
        char *end = nullptr;
        e->VAF[i] = std::strtod(f.c_str(), &end);
        if (end == f.c_str()) formatError("VAF value " + f, line);
        if (prm.INPUT_FORMAT == Parameters::SNV_WITH_PROFILE) continue;
        if (e->VAF[i] < prm.MIN_VAF_PRESENT) {
            if (e->VAF[i] >= prm.MAX_VAF_ABSENT) { e->isRobust = false; e->ambigPresenceProfile += "*"; }
            else e->ambigPresenceProfile += "0";
            e->presenceProfile += "0";
        } else {
            e->presenceProfile += "1";
            e->ambigPresenceProfile += "1";
        }
    }
    return e;
}
void SNVDataStore::processSNVEntry(SNVEntry *e, int normalSample) {                            // :385-409
    const int S = getNumSamples();
    if (prm.INPUT_FORMAT != Parameters::SNV_WITH_PROFILE && e->presenceProfile[normalSample] == '1') return;
    for (int i = 0; i < S; i++) if (e->VAF[i] > prm.MAX_ALLOWED_VAF) return;
    if (e->isRobust && e->presenceProfile.find('1') == std::string::npos) return;
    somaticSNVs.push_back(e);
    if (e->isRobust) { put(e->presenceProfile); tag2SNVs[e->presenceProfile].push_back(e); }
    else ambiguousSNVs.push_back(e);
}
static std::vector<std::string> readLines(const std::string &f) {
    std::ifstream in(f);
    if (!in) formatError("Could not read file: " + f, "");
    std::vector<std::string> lines;
    std::string l;
    while (std::getline(in, l)) { if (!l.empty() && l.back() == '\r') l.pop_back(); lines.push_back(l); }
    return lines;
}
void SNVDataStore::loadSNVFile(const std::string &f, int normalSample) {                       // :522-545
    auto lines = readLines(f);
    if (lines.empty()) formatError("Empty file", "");
    parseHeader(lines[0]);
    for (size_t i = 1; i < lines.size(); i++) {
        allSNVs.push_back(parseSNVEntry(lines[i], (int)i));
        processSNVEntry(allSNVs.back().get(), normalSample);
    }
}
void SNVDataStore::loadSNVFileWithClusters(const std::string &f, const std::string &cf) {      // :547-621
    auto lines = readLines(f);
    if (lines.empty()) formatError("Empty file", "");
    parseHeader(lines[0]);
    for (size_t i = 1; i < lines.size(); i++) {
        allSNVs.push_back(parseSNVEntry(lines[i], (int)i));
        somaticSNVs.push_back(allSNVs.back().get());     // no filtering
    }
    auto cl = readLines(cf);
    if (cl.empty()) formatError("Empty file", "");
    const int S = getNumSamples();
    int numClusters = 0;
    for (const std::string &line : cl) {
        auto parts = split(line, '\t');
        if ((int)parts.size() != 2 + S)
            formatError("Expecting 2 fields and " + std::to_string(S) + " samples based on the SSNV file", line);
        std::string profile = trim(parts[0]);
        if ((int)profile.size() != S)
            formatError("Cluster presence profile " + profile + " length does not match the number of input samples", line);
        auto c = std::make_shared<Cluster>();
        for (int i = 0; i < S; i++) {
            char *end = nullptr;
            double v = std::strtod(parts[i + 1].c_str(), &end);
            if (end == parts[i + 1].c_str()) formatError("Cluster centroid VAF value " + parts[i + 1], line);
            if (profile[i] == '1') c->centroid.push_back(v);
        }
        c->id = numClusters++;
        c->stdDev.assign(S, 0.0);
        put(profile);
        const int startId = (int)tag2SNVs[profile].size();
        int k = 0;
        for (const std::string &m : split(parts[S + 1], ',')) {
            std::string t = trim(m);
            char *end = nullptr;
            long id = std::strtol(t.c_str(), &end, 10);
            if (t.empty() || *end) formatError("SSNV member entry " + t + " is not a valid number ", line);
            if (id < 1 || id > (long)somaticSNVs.size()) formatError("SSNV member entry " + t + " is out of range", line);
            c->members.push_back(startId + k++);
            SNVEntry *snv = somaticSNVs[id - 1];
            snv->presenceProfile = profile;
            snv->isRobust = true;
            tag2SNVs[profile].push_back(snv);
        }
        tag2Clusters[profile].push_back(c);
    }
}
bool SNVDataStore::canConvert(const SNVEntry &s, const std::string &target) const {            // :316-329
    for (size_t i = 0; i < s.presenceProfile.size(); i++)
        if (s.presenceProfile[i] != target[i]) {
            if (s.presenceProfile[i] == '1') return false;
            if (!s.evidenceOfPresence((int)i, prm)) return false;
        }
    return true;
}
void SNVDataStore::assignAmbiguousSNVs() {                                                     // :113-191, 197-267
    if (ambiguousSNVs.empty()) return;
    const int S = getNumSamples();
    const std::string all1(S, '1'), all0(S, '0');
    std::vector<std::string> targets = groupTags();
    if (!tag2SNVs.count(all0)) targets.push_back(all0);
    targets.push_back(all1);
    std::vector<SNVEntry *> toRemove;
    auto removed = [&](SNVEntry *e) { return std::find(toRemove.begin(), toRemove.end(), e) != toRemove.end(); };
    for (SNVEntry *snv : ambiguousSNVs) {
        double best = 0;
        std::string bestT;
        bool germline = false;
        for (const std::string &target : targets) {
            if (!canConvert(*snv, target)) continue;
            if (target == all1) { toRemove.push_back(snv); germline = true; break; }
            if (target == all0) {
                double d = 0;
                for (int i = 0; i < S; i++)
                    if (snv->presenceProfile[i] == '1' || snv->evidenceOfPresence(i, prm)) d += 0.0001 / snv->VAF[i];
                if (d > best) { best = d; bestT = target; }
                continue;
            }
            for (SNVEntry *ts : tag2SNVs[target]) {
                double d = 0;
                for (int i = 0; i < S; i++) {
                    if (ts->presenceProfile[i] == '0' && snv->presenceProfile[i] == '0' && !snv->evidenceOfPresence(i, prm)) continue;
                    double mn = std::min(ts->VAF[i], snv->VAF[i]), mx = std::max(ts->VAF[i], snv->VAF[i]);
                    if (mn == 0) mn = 0.0001;
                    d += mn / mx;
                }
                if (d > best) { best = d; bestT = target; }
            }
        }
        (void)germline;
        if (bestT == all0) { snv->presenceProfile = bestT; continue; }
        if (best != 0 && !bestT.empty() && best >= (double)std::count(bestT.begin(), bestT.end(), '1') * prm.MIN_VAF_TARGET_RATIO_PER_SAMPLE) {
            snv->presenceProfile = bestT;
            if (!removed(snv)) toRemove.push_back(snv);
            tag2SNVs[bestT].push_back(snv);
            continue;
        }
    }
    std::vector<SNVEntry *> rest;
    for (SNVEntry *e : ambiguousSNVs) if (!removed(e)) rest.push_back(e);
    ambiguousSNVs = rest;
    // mergeAmbiguousSNVs: greedy set cover over all reachable target profiles
    std::vector<std::string> ainsert, ginsert;
    std::map<std::string, std::vector<SNVEntry *>> adj, groups;
    std::function<void(const std::string &, SNVEntry *, std::vector<std::string> &)> extend =
        [&](const std::string &partial, SNVEntry *snv, std::vector<std::string> &out) {
            const int k = (int)partial.size();
            if (k == S) { out.push_back(partial); return; }
            extend(partial + snv->presenceProfile[k], snv, out);
            if (snv->presenceProfile[k] == '0' && snv->evidenceOfPresence(k, prm)) extend(partial + '1', snv, out);
        };
    std::vector<std::string> tl;
    for (SNVEntry *snv : rest) extend("", snv, tl);
    for (auto &t : tl) if (!adj.count(t)) { adj[t]; ainsert.push_back(t); }
    const size_t adj_hwm = ainsert.size();
    auto ordered = [&](const std::vector<std::string> &keys, size_t hw) {
        std::vector<int32_t> h;
        for (auto &t : keys) h.push_back(javaStringHash(t));
        std::vector<std::string> out;
        for (size_t i : javaHashOrder(h, hw)) out.push_back(keys[i]);
        return out;
    };
    for (SNVEntry *snv : rest) for (auto &t : ordered(ainsert, adj_hwm)) if (canConvert(*snv, t)) adj[t].push_back(snv);
    while (!rest.empty()) {
        size_t mx = 0;
        std::string mset;
        std::vector<std::string> live;
        for (auto &t : ainsert) if (adj.count(t)) live.push_back(t);
        for (auto &t : ordered(live, adj_hwm)) if (adj[t].size() > mx) { mx = adj[t].size(); mset = t; }
        if (mx <= 1) break;
        groups[mset] = adj[mset];
        ginsert.push_back(mset);
        for (SNVEntry *entry : std::vector<SNVEntry *>(adj[mset])) {
            entry->presenceProfile = mset;
            for (auto &kv : adj) {
                if (kv.first == mset) continue;
                auto &l = kv.second;
                auto it = std::find(l.begin(), l.end(), entry);
                if (it != l.end()) l.erase(it);
            }
            auto it = std::find(rest.begin(), rest.end(), entry);
            if (it != rest.end()) rest.erase(it);
        }
        adj.erase(mset);
    }
    for (SNVEntry *snv : rest) {
        std::string tag;
        for (int i = 0; i < S; i++) {
            if (snv->presenceProfile[i] == '0' && snv->evidenceOfPresence(i, prm)) {
                double d0 = snv->VAF[i] - prm.MAX_VAF_ABSENT, d1 = prm.MIN_VAF_PRESENT - snv->VAF[i];
                tag += d1 < d0 ? '1' : snv->presenceProfile[i];
            } else tag += snv->presenceProfile[i];
        }
        if (!groups.count(tag)) { groups[tag]; ginsert.push_back(tag); }
        groups[tag].push_back(snv);
    }
    for (auto &t : ordered(ginsert, ginsert.size())) {
        if (t == all0) continue;
        if (!tag2SNVs.count(t)) { put(t); tag2SNVs[t] = groups[t]; }
    }
    if (tag2SNVs.count(all0)) erase(all0);
}
SNVDataStore::SNVDataStore(const std::string &snvFile, const std::string &clustersFile, int normalSample, const Parameters &p)
    : prm(p) {                                                                                  // :67-107
    if (clustersFile.empty()) loadSNVFile(snvFile, normalSample);
    else loadSNVFileWithClusters(snvFile, clustersFile);
    if (prm.INPUT_FORMAT == Parameters::SNV_WITH_PROFILE || !clustersFile.empty()) return;
    std::vector<std::string> small;
    for (auto &t : groupTags())
        if ((int)tag2SNVs[t].size() < prm.MIN_GROUP_PROFILE_SUPPORT) {
            for (SNVEntry *e : tag2SNVs[t]) ambiguousSNVs.push_back(e);
            small.push_back(t);
        }
    for (auto &t : small) erase(t);
    assignAmbiguousSNVs();
    // filterGroups (:347-374)
    std::vector<std::string> out;
    for (auto &t : groupTags()) {
        if (t.find('1') == std::string::npos || (int)tag2SNVs[t].size() < prm.MIN_SNVS_PER_GROUP) { out.push_back(t); continue; }
        if (prm.MIN_ROBUST_SNVS_PER_GROUP > 0) {
            int nr = 0;
            for (SNVEntry *e : tag2SNVs[t]) if (e->isRobust) nr++;
            if (nr < prm.MIN_ROBUST_SNVS_PER_GROUP) out.push_back(t);
        }
    }
    for (auto &t : out) erase(t);
}

// ---------------------------------------------------------------------------------------------
// PHYNetwork
// ---------------------------------------------------------------------------------------------
double PHYNetwork::getAAF(int n, int s) const {                                                // PHYNode.java:161-173
    const PHYNode &nd = nodesById[n];
    if (nd.isRoot) return prm.VAF_MAX;
    int k = nd.snvGroup->getSampleIndex(s);
    return k < 0 ? 0.0 : nd.cluster->centroid[k];
}
double PHYNetwork::getStdDev(int n, int s) const {                                             // PHYNode.java:179-192
    const PHYNode &nd = nodesById[n];
    if (nd.isRoot) return 0.0;
    int k = nd.snvGroup->getSampleIndex(s);
    if (k < 0 || nd.cluster->stdDev.empty()) return 0.0;
    return nd.cluster->stdDev[k];
}
double PHYNetwork::margin(int from, int to, int i) const {                                     // :237-263
    auto se = [&](int n) {
        if (nodesById[n].isRoot) return prm.VAF_ERROR_MARGIN;
        return 1.96 * getStdDev(n, i) / std::sqrt((double)nodesById[n].cluster->members.size());
    };
    double s = se(from) + se(to);
    return s > prm.VAF_ERROR_MARGIN ? s : prm.VAF_ERROR_MARGIN;
}
void PHYNetwork::addEdge(int a, int b) {                                                       // :277-286
    if (std::find(edges[a].begin(), edges[a].end(), b) == edges[a].end()) { edges[a].push_back(b); numEdges++; }
}
int PHYNetwork::checkAndAddEdge(int n1, int n2) {                                              // :185-232
    const int S = numSamples;
    int c12 = 0, c21 = 0;
    double e12 = 0, e21 = 0;
    for (int i = 0; i < S; i++) {
        const double A = getAAF(n1, i), B = getAAF(n2, i);
        if (A == 0 && B != 0) break;
        c12 += A >= (B - margin(n1, n2, i)) ? 1 : 0;
        if (A < B) e12 += B - A;
    }
    for (int i = 0; i < S; i++) {
        const double A = getAAF(n1, i), B = getAAF(n2, i);
        if (B == 0 && A != 0) break;
        c21 += B >= (A - margin(n2, n1, i)) ? 1 : 0;
        if (B < A) e21 += A - B;
    }
    if (c12 == S) {
        if (c21 == S) {
            if (e12 < e21) { addEdge(n1, n2); return 0; }
            addEdge(n2, n1);
            return 1;
        }
        addEdge(n1, n2);
        return 0;
    } else if (c21 == S) {
        addEdge(n2, n1);
        return 1;
    }
    return -1;
}
PHYNetwork::PHYNetwork(std::vector<SNVGroup *> groups, int S, const Parameters &p) : prm(p), groups_(groups) {  // :94-175
    numSamples = S;
    auto addNode = [&](PHYNode nd, int level) {
        nd.nodeId = numNodes;
        nd.level = level;
        nodes[level].push_back(nd.nodeId);
        nodesById.push_back(nd);
        edges.emplace_back();
        numNodes++;
        return nd.nodeId;
    };
    PHYNode root;
    root.isRoot = true;
    addNode(root, S + 1);
    for (SNVGroup *g : groups) {
        std::vector<int> ids;
        for (auto &c : g->subPopulations) {
            PHYNode nd;
            nd.snvGroup = g;
            nd.cluster = c;
            ids.push_back(addNode(nd, g->getNumSamples()));
        }
        for (size_t i = 0; i < ids.size(); i++) for (size_t j = i + 1; j < ids.size(); j++) checkAndAddEdge(ids[i], ids[j]);
    }
    auto level = [&](int l) -> const std::vector<int> * { auto it = nodes.find(l); return it == nodes.end() ? nullptr : &it->second; };
    for (int i = S + 1; i > 0; i--) {
        auto fr = level(i);
        if (!fr) continue;
        int j = i - 1;
        auto to = level(j);
        while (!to && j > 0) { j--; to = level(j); }
        if (!to) continue;
        for (int n1 : *fr) for (int n2 : *to) checkAndAddEdge(n1, n2);
    }
    if (prm.ALL_EDGES) {                                                                       // addAllHiddenEdges :302-316
        for (int i = S + 1; i > 0; i--) {
            auto fr = level(i);
            if (!fr) continue;
            for (int j = i - 1; j >= 1; j--) {
                auto to = level(j);
                if (!to) continue;
                for (int n1 : *fr) for (int n2 : *to) checkAndAddEdge(n1, n2);
            }
        }
    }
    std::vector<int> mask(numNodes, 0);
    for (int a = 0; a < numNodes; a++) for (int b : edges[a]) mask[b] = 1;
    for (int i = 1; i < numNodes; i++) {
        if (mask[i]) continue;
        bool found = false;
        for (int j = nodesById[i].level + 2; j <= S + 1 && !found; j++) {
            auto fr = level(j);
            if (!fr) continue;
            for (int n2 : *fr) if (checkAndAddEdge(n2, i) == 0) { found = true; break; }
        }
        if (!found) addEdge(0, i);
    }
}

std::vector<PHYTree> PHYNetwork::getLineageTrees(int numSave, lichee_search_stats *stats_out, int device) const {
    std::vector<int32_t> off(numNodes + 1, 0), idx;
    for (int i = 0; i < numNodes; i++) { for (int c : edges[i]) idx.push_back(c); off[i + 1] = (int32_t)idx.size(); }
    if (idx.empty()) idx.push_back(0);
    std::vector<double> vaf((size_t)numNodes * numSamples);
    for (int i = 0; i < numNodes; i++) for (int s = 0; s < numSamples; s++) vaf[(size_t)i * numSamples + s] = getAAF(i, s);
    lichee_search_params prmc;
    lichee_default_params(&prmc);
    prmc.vaf_error_margin = prm.VAF_ERROR_MARGIN;
    prmc.num_save = std::max(1, numSave);
    if (prmc.num_save > LICHEE_MAX_SAVE) {                 // the library ranks at most LICHEE_MAX_SAVE trees on the device
        std::cerr << "lichee_b200: -s " << numSave << " exceeds " << LICHEE_MAX_SAVE << "; the first " << LICHEE_MAX_SAVE
                  << " trees of the ranked list are saved\n";
        prmc.num_save = LICHEE_MAX_SAVE;
    }
    prmc.max_num_trees = prm.MAX_NUM_TREES;
    prmc.max_num_grow_calls = prm.MAX_NUM_GROW_CALLS;
    prmc.device = device;
    lichee_search_stats st;
    std::vector<double> scores(prmc.num_save);
    std::vector<int32_t> parents((size_t)prmc.num_save * numNodes), orders((size_t)prmc.num_save * numNodes);
    int rc = lichee_get_lineage_trees(numNodes, numSamples, off.data(), idx.data(), vaf.data(), &prmc, &st, scores.data(),
                                      parents.data(), orders.data());
    if (rc != LICHEE_OK) throw std::runtime_error(std::string("lichee_b200: ") + lichee_last_error());
    if (stats_out) *stats_out = st;
    std::vector<PHYTree> out;
    for (uint32_t k = 0; k < st.num_saved; k++) {
        PHYTree t;
        for (int j = 0; j < numNodes; j++) {
            int v = orders[(size_t)k * numNodes + j];
            if (v < 0) break;
            t.treeNodes.push_back(v);
            int p = parents[(size_t)k * numNodes + v];
            if (p >= 0) {
                if (!t.treeEdges.count(p)) t.keyOrder.push_back(p);
                t.treeEdges[p].push_back(v);
            }
        }
        t.errorScore = scores[k];
        out.push_back(t);
    }
    return out;
}

bool PHYNetwork::fixNetwork(std::vector<SNVGroup *> &groups, std::unique_ptr<PHYNetwork> &out) const {  // :325-358
    const Cluster *toRemove = nullptr;
    SNVGroup *group = nullptr;
    for (const PHYNode &n : nodesById) {        // HashMap<Integer,PHYNode>.values(): ascending id
        if (!n.snvGroup) continue;
        const Cluster *c = n.cluster.get();
        if (!c->robust && (!toRemove || c->members.size() < toRemove->members.size())) { toRemove = c; group = n.snvGroup; }
    }
    if (toRemove) group->removeCluster(toRemove);
    // The reference rebuilds from a HashSet<SNVGroup> (identity hash: iteration order is not
    // deterministic in the JVM); the group order of the current network is kept here.
    out = std::make_unique<PHYNetwork>(groups, numSamples, prm);
    return toRemove != nullptr;
}

std::string PHYNetwork::toString() const {                                                     // :787-812
    std::string g = "--- PHYLOGENETIC CONSTRAINT GRAPH --- \n";
    g += "numNodes = " + std::to_string(numNodes) + ", numEdges = " + std::to_string(numEdges) + "\n";
    g += "EDGES: \n";
    for (int a = 0; a < numNodes; a++) for (int b : edges[a]) g += std::to_string(a) + " -> " + std::to_string(b) + "\n";
    return g;
}
std::string PHYNetwork::getNodesWithMembersAsString() const {                                  // :845-869
    std::string s;
    for (int i = numSamples + 1; i >= 0; i--) {
        auto it = nodes.find(i);
        if (it == nodes.end()) continue;
        for (int id : it->second) {
            const PHYNode &n = nodesById[id];
            if (n.isRoot) continue;
            s += std::to_string(id) + "\t" + n.snvGroup->tag + "\t[";
            for (double c : n.cluster->centroid) s += " " + decimalFormat(c, 2);
            s += "]";
            for (int m : n.cluster->members) s += "\tsnv" + std::to_string(n.snvGroup->snvs[m]->id);
            s += "\n";
        }
    }
    return s;
}
std::string PHYNetwork::getNodeMembersOnlyAsString() const {                                   // :871-886
    std::string s;
    for (int i = numSamples + 1; i >= 0; i--) {
        auto it = nodes.find(i);
        if (it == nodes.end()) continue;
        for (int id : it->second) {
            const PHYNode &n = nodesById[id];
            if (n.isRoot) continue;
            for (int m : n.cluster->members) {
                const SNVEntry *e = n.snvGroup->snvs[m];
                s += "snv" + std::to_string(e->id) + ": " + std::to_string(e->chr) + " " + std::to_string(e->position) + " " + e->description + "\n";
            }
        }
    }
    return s;
}
std::string PHYNetwork::treeToString(const PHYTree &t) const {                                 // PHYTree.java:176-185
    // treeEdges.keySet(): HashMap<PHYNode,..> with hashCode() == nodeId
    std::vector<int32_t> h(t.keyOrder.begin(), t.keyOrder.end());
    std::string g;
    for (size_t i : javaHashOrder(h, h.size())) {
        int p = t.keyOrder[i];
        for (int c : t.treeEdges.at(p)) g += std::to_string(p) + " -> " + std::to_string(c) + "\n";
    }
    return g;
}
void PHYNetwork::lineageHelper(const PHYTree &t, std::string &out, std::string indent, int n, int sampleId) const {  // :256-269
    indent += ".....";
    const PHYNode &nd = nodesById[n];
    if (nd.snvGroup->containsSample(sampleId))
        out += indent + nd.snvGroup->tag + ": " + decimalFormat(getAAF(n, sampleId), 3) + " [" + decimalFormat(getStdDev(n, sampleId), 3) + "]\n";
    auto it = t.treeEdges.find(n);
    if (it != t.treeEdges.end()) for (int c : it->second) lineageHelper(t, out, indent, c, sampleId);
}
std::string PHYNetwork::getLineage(const PHYTree &t, int sampleId, const std::string &name) const {  // :242-254
    std::string s = "\tSample lineage decomposition: " + name + "\nGL\n";
    auto it = t.treeEdges.find(t.treeNodes.empty() ? 0 : t.treeNodes[0]);
    if (it != t.treeEdges.end()) for (int c : it->second) lineageHelper(t, s, "", c, sampleId);
    return s;
}

// ---------------------------------------------------------------------------------------------
// LineageEngine.buildLineage
// ---------------------------------------------------------------------------------------------
int buildLineage(const Args &args, Parameters prm) {
    SNVDataStore db(args.inputFileName, args.clustersFileName, args.normalSampleId, prm);          // step 1
    std::vector<std::unique_ptr<SNVGroup>> owned;                                                   // step 2
    for (const std::string &tag : db.groupTags())
        owned.push_back(std::make_unique<SNVGroup>(tag, db.tag2SNVs[tag], db.isRobustGroup(tag)));
    if (owned.empty()) { std::cerr << "All SNV groups have been filtered out.\n"; return 0; }
    for (auto &g : owned) {                                                                          // step 3
        if (!args.clustersFileName.empty()) {
            g->subPopulations = db.tag2Clusters[g->tag];
        } else if (args.singleCluster) {
            auto c = std::make_shared<Cluster>();
            c->id = 0;
            for (size_t i = 0; i < g->snvs.size(); i++) c->members.push_back((int)i);
            c->recomputeCentroidAndStdDev(g->alleleFreqBySample, g->getNumSamples());
            g->setSubPopulations({c}, prm);
        } else {
            throw std::runtime_error("SSNV clustering uses Weka EM in the reference (AAFClusterer.em), which is not part of this "
                                     "library: pass -clustersFile <file>, or -singleCluster for one cluster per SSNV group");
        }
    }
    std::vector<SNVGroup *> groups;
    for (auto &g : owned) groups.push_back(g.get());
    auto net = std::make_unique<PHYNetwork>(groups, db.getNumSamples(), prm);                       // step 4
    if (!args.dumpNetworkFileName.empty()) {
        std::ofstream d(args.dumpNetworkFileName);
        d.precision(17);
        d << net->numNodes << " " << net->numSamples << "\n";
        for (int i = 0; i < net->numNodes; i++) {
            d << (net->nodesById[i].isRoot ? std::string("GL") : net->nodesById[i].snvGroup->tag) << " "
              << (net->nodesById[i].isRoot ? 0 : (int)net->nodesById[i].cluster->members.size());
            for (int s = 0; s < net->numSamples; s++) d << " " << net->getAAF(i, s);
            d << " |";
            for (int c : net->edges[i]) d << " " << c;
            d << "\n";
        }
        return 0;
    }
    const int numSave = std::max(1, args.numSave);
    lichee_search_stats st;
    std::vector<PHYTree> trees = net->getLineageTrees(numSave, &st, args.device);                   // step 5 + 6
    std::cerr << "Found " << st.num_trees << " valid tree(s)\r\n";
    if (st.num_trees == 0) {
        std::cerr << "Adjusting the network...\r\n";
        int delta = 0;
        do {
            const int numNodes = net->numNodes;
            std::unique_ptr<PHYNetwork> fixed;
            net->fixNetwork(groups, fixed);
            net = std::move(fixed);
            trees = net->getLineageTrees(numSave, &st, args.device);
            delta = numNodes - net->numNodes;
        } while (delta != 0 && st.num_trees <= 0);
        if (st.num_trees <= 0) {
            prm.ALL_EDGES = true;
            net = std::make_unique<PHYNetwork>(groups, db.getNumSamples(), prm);
            trees = net->getLineageTrees(numSave, &st, args.device);
        }
        std::cerr << "Found " << st.num_trees << " valid trees after network adjustments\r\n";
    }
    if (!trees.empty() && args.numSave > 0) {                                                       // step 8, writeTreesToTxtFile :161-189
        std::ofstream fw(args.outputFileName);
        if (!fw) { std::cerr << "Failed to write to the file: " << args.outputFileName << "\n"; return -1; }
        fw << "Nodes:\n" << net->getNodesWithMembersAsString() << "\n";
        for (int i = 0; i < args.numSave && i < (int)trees.size(); i++) {
            fw << "****Tree " << i << "****\n" << net->treeToString(trees[i]);
            fw << "Error score: " << javaDoubleToString(trees[i].getErrorScore()) << "\n\n";
            fw << "Sample decomposition: \n";
            for (int j = 0; j < db.getNumSamples(); j++) fw << net->getLineage(trees[i], j, db.sampleNames[j]) << "\n";
            fw << "\n";
        }
        fw << "SNV info:\n" << net->getNodeMembersOnlyAsString() << "\n";
    }
    return 0;
}

}  // namespace lichee
