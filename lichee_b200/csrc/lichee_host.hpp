// lichee_host.hpp -- C++ host mirror of the reference code either side of the tree search.
//
// The reference is compiled code (Java) and this image has no JVM, so the host above the C ABI is
// C++; class and method names follow LICHeE/src/lineage so the call sites read like the reference:
//   Parameters            Parameters.java
//   SNVEntry              SNVEntry.java
//   SNVDataStore          SNVDataStore.java   (loadSNVFile, loadSNVFileWithClusters, calling, groups)
//   Cluster / SNVGroup    AAFClusterer.Cluster, SNVGroup.java (setSubPopulations, removeCluster)
//   PHYNode / PHYNetwork  PHYNode.java, PHYNetwork.java (constructor, checkAndAddEdge,
//                         getAAFErrorMargin, addAllHiddenEdges, fixNetwork, getLineageTrees ->
//                         libLICHeE_b200.so)
//   PHYTree               PHYTree.java (toString, getLineage, getErrorScore)
//   LineageEngine         LineageEngine.java (buildLineage steps 1-8 without GUI/DOT,
//                         writeTreesToTxtFile)
// Not mirrored: AAFClusterer.em (Weka EM; lib/weka.jar is absent from the reference checkout).  The
// CLI takes the reference's own -clustersFile, or -singleCluster (one cluster per SSNV group).
#ifndef LICHEE_HOST_HPP
#define LICHEE_HOST_HPP

#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/lichee_b200.h"

namespace lichee {

struct Parameters {                       // Parameters.java defaults
    enum Format { SNV, SNV_WITH_PROFILE };
    Format INPUT_FORMAT = SNV;
    double MAX_VAF_ABSENT = 0.03, MIN_VAF_PRESENT = 0.05, MAX_ALLOWED_VAF = 0.6;
    int MIN_GROUP_PROFILE_SUPPORT = 2, MIN_SNVS_PER_GROUP = 1, MIN_ROBUST_SNVS_PER_GROUP = 0;
    double MIN_VAF_TARGET_RATIO_PER_SAMPLE = 0.5;
    int MIN_CLUSTER_SIZE = 2, MIN_PRIVATE_CLUSTER_SIZE = 1, MIN_ROBUST_CLUSTER_SUPPORT = 2;
    double MAX_COLLAPSE_CLUSTER_DIFF = 0.2;
    double VAF_MAX = 0.5, VAF_ERROR_MARGIN = 0.1;
    bool ALL_EDGES = false, CP = false;
    int64_t MAX_NUM_TREES = 100000, MAX_NUM_GROW_CALLS = 100000000;
};

// Iteration order of a java.util.HashMap key set: buckets by index, insertion order inside a
// bucket; the table doubles when size exceeds 0.75*capacity and never shrinks.
std::vector<size_t> javaHashOrder(const std::vector<int32_t> &hashes, size_t max_size);
int32_t javaStringHash(const std::string &s);
std::string javaDoubleToString(double d);            // Double.toString
std::string decimalFormat(double d, int max_frac);   // DecimalFormat("#.##") / ("#.###"), HALF_EVEN

struct SNVEntry {
    int id = 0, chr = 0, position = 0;
    std::string description, presenceProfile, ambigPresenceProfile, line;
    std::vector<double> VAF;
    bool isRobust = true;
    bool evidenceOfPresence(int s, const Parameters &p) const { return VAF[s] > p.MAX_VAF_ABSENT; }
};

struct Cluster {
    int id = 0;
    std::vector<double> centroid, stdDev;
    std::vector<int> members;
    bool robust = false;
    void recomputeCentroidAndStdDev(const std::vector<std::vector<double>> &data, int numFeatures);
};

struct SNVGroup {
    std::string tag;
    std::vector<int> sampleIndex;                      // samples with a '1' in the tag
    std::vector<SNVEntry *> snvs;
    std::vector<std::vector<double>> alleleFreqBySample;
    std::vector<std::shared_ptr<Cluster>> subPopulations;
    bool isRobust = false;
    SNVGroup(const std::string &tag, const std::vector<SNVEntry *> &snvs, bool robust);
    int getNumSamples() const { return (int)sampleIndex.size(); }
    int getSampleIndex(int sampleId) const;
    bool containsSample(int sampleId) const { return getSampleIndex(sampleId) >= 0; }
    void setSubPopulations(std::vector<std::shared_ptr<Cluster>> clusters, const Parameters &p);
    void removeCluster(const Cluster *c);
};

class SNVDataStore {
public:
    SNVDataStore(const std::string &snvFile, const std::string &clustersFile, int normalSample, const Parameters &p);
    int getNumSamples() const { return (int)sampleNames.size(); }
    std::vector<std::string> sampleNames;
    std::vector<std::unique_ptr<SNVEntry>> allSNVs;
    std::vector<SNVEntry *> somaticSNVs, ambiguousSNVs;
    // tag2SNVs with HashMap iteration order
    std::vector<std::string> groupTags() const;
    std::map<std::string, std::vector<SNVEntry *>> tag2SNVs;
    std::map<std::string, std::vector<std::shared_ptr<Cluster>>> tag2Clusters;
    bool isRobustGroup(const std::string &tag) const;
private:
    Parameters prm;
    std::vector<std::string> insertion;                // key insertion order of tag2SNVs
    size_t hwm = 0;                                    // largest size tag2SNVs ever had
    void put(const std::string &tag);
    void erase(const std::string &tag);
    void parseHeader(const std::string &line);
    std::unique_ptr<SNVEntry> parseSNVEntry(const std::string &line, int lineId);
    void processSNVEntry(SNVEntry *e, int normalSample);
    void loadSNVFile(const std::string &f, int normalSample);
    void loadSNVFileWithClusters(const std::string &f, const std::string &cf);
    void assignAmbiguousSNVs();
    bool canConvert(const SNVEntry &s, const std::string &target) const;
};

struct PHYNode {
    int nodeId = 0, level = 0;
    bool isRoot = false;
    SNVGroup *snvGroup = nullptr;
    std::shared_ptr<Cluster> cluster;
};

struct PHYTree {
    std::vector<int> treeNodes;                         // node ids, insertion order
    std::map<int, std::vector<int>> treeEdges;          // parent -> children in insertion order
    std::vector<int> keyOrder;                          // parents in first-insertion order
    double errorScore = -1;
    double getErrorScore() const { return errorScore; }
};

class PHYNetwork {
public:
    PHYNetwork(std::vector<SNVGroup *> groups, int totalNumSamples, const Parameters &p);
    std::vector<PHYNode> nodesById;
    std::map<int, std::vector<int>> nodes;              // level -> node ids
    std::vector<std::vector<int>> edges;                // adjacency lists, insertion order
    int numNodes = 0, numEdges = 0, numSamples = 0;
    double getAAF(int node, int sample) const;
    double getStdDev(int node, int sample) const;
    // getLineageTrees() + evaluateLineageTrees() on the GPU; throws std::runtime_error on failure
    std::vector<PHYTree> getLineageTrees(int numSave, lichee_search_stats *stats = nullptr, int device = 0) const;
    // fixNetwork(): drops the smallest non-robust cluster and rebuilds; returns false if nothing to drop
    bool fixNetwork(std::vector<SNVGroup *> &groups, std::unique_ptr<PHYNetwork> &out) const;
    std::string toString() const;
    std::string getNodesWithMembersAsString() const;
    std::string getNodeMembersOnlyAsString() const;
    std::string treeToString(const PHYTree &t) const;                       // PHYTree.toString
    std::string getLineage(const PHYTree &t, int sampleId, const std::string &sampleName) const;
    const Parameters prm;
private:
    std::vector<SNVGroup *> groups_;
    int checkAndAddEdge(int n1, int n2);
    void addEdge(int a, int b);
    double margin(int from, int to, int i) const;
    void lineageHelper(const PHYTree &t, std::string &out, std::string indent, int n, int sampleId) const;
};

struct Args {
    std::string inputFileName, outputFileName, clustersFileName, dumpNetworkFileName;
    int normalSampleId = 0, numSave = 1, device = 0;
    bool singleCluster = false, verbose = false;
};

// LineageEngine.buildLineage (LineageEngine.java:69-157) without GUI / DOT export.
int buildLineage(const Args &args, Parameters prm);

}  // namespace lichee
#endif
