// lichee_cli.cpp -- `lichee_b200_build`: the reference's `lichee -build` command line
// (LineageEngine.main, LineageEngine.java:207-402) on top of libLICHeE_b200.so.
// GUI / DOT options are accepted and ignored; clustering needs -clustersFile or -singleCluster.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "lichee_host.hpp"

static void usage() {
    std::cerr <<
        "usage: lichee_b200_build -build -i <file> [-o <file>] [-n <normal sample id>] [-clustersFile <file> | -singleCluster]\n"
        "       [-sampleProfile] [-cp] [-s <trees to save>] [-maxVAFAbsent x] [-minVAFPresent x] [-maxVAFValid x]\n"
        "       [-minProfileSupport n] [-minClusterSize n] [-minPrivateClusterSize n] [-minRobustNodeSupport n]\n"
        "       [-maxClusterDist x] [-c] [-e x] [-v]   (reference options, LineageEngine.java:208-242)\n"
        "       [-dumpNetwork <file>] [-device n] [-maxNumTrees n] [-maxNumGrowCalls n]   (additions)\n";
}

int main(int argc, char **argv) {
    lichee::Args a;
    lichee::Parameters p;
    bool build = false, haveN = false, haveAbsent = false, havePresent = false;
    auto need = [&](int &i) -> const char * {
        if (i + 1 >= argc) { std::cerr << "Missing argument for option: " << argv[i] << "\n"; usage(); std::exit(-1); }
        return argv[++i];
    };
    for (int i = 1; i < argc; i++) {
        std::string o = argv[i];
        while (o.size() > 1 && o[0] == '-' && o[1] == '-') o.erase(0, 1);
        if (o == "-build") build = true;
        else if (o == "-i") a.inputFileName = need(i);
        else if (o == "-o") a.outputFileName = need(i);
        else if (o == "-cp") p.CP = true;
        else if (o == "-sampleProfile") p.INPUT_FORMAT = lichee::Parameters::SNV_WITH_PROFILE;
        else if (o == "-n" || o == "-normal") { a.normalSampleId = std::atoi(need(i)); haveN = true; }
        else if (o == "-clustersFile") a.clustersFileName = need(i);
        else if (o == "-s" || o == "-save") a.numSave = std::atoi(need(i));
        else if (o == "-maxVAFAbsent" || o == "-absent") { p.MAX_VAF_ABSENT = std::atof(need(i)); haveAbsent = true; }
        else if (o == "-minVAFPresent" || o == "-present") { p.MIN_VAF_PRESENT = std::atof(need(i)); havePresent = true; }
        else if (o == "-maxVAFValid") p.MAX_ALLOWED_VAF = std::atof(need(i));
        else if (o == "-minProfileSupport") p.MIN_GROUP_PROFILE_SUPPORT = std::atoi(need(i));
        else if (o == "-minClusterSize") p.MIN_CLUSTER_SIZE = std::atoi(need(i));
        else if (o == "-minPrivateClusterSize") p.MIN_PRIVATE_CLUSTER_SIZE = std::atoi(need(i));
        else if (o == "-minRobustNodeSupport") p.MIN_ROBUST_CLUSTER_SUPPORT = std::atoi(need(i));
        else if (o == "-maxClusterDist") p.MAX_COLLAPSE_CLUSTER_DIFF = std::atof(need(i));
        else if (o == "-c" || o == "-completeNetwork") p.ALL_EDGES = true;
        else if (o == "-e") p.VAF_ERROR_MARGIN = std::atof(need(i));
        else if (o == "-nTreeQPCheck") need(i);                       // QP check is off in the reference by default; not mirrored
        else if (o == "-v" || o == "-verbose") a.verbose = true;
        else if (o == "-showTree" || o == "-tree" || o == "-dotFile") need(i);   // GUI / DOT: not on this path
        else if (o == "-showNetwork" || o == "-net" || o == "-color" || o == "-dot") {}
        else if (o == "-singleCluster") a.singleCluster = true;
        else if (o == "-dumpNetwork") a.dumpNetworkFileName = need(i);
        else if (o == "-device") a.device = std::atoi(need(i));
        else if (o == "-maxNumTrees") p.MAX_NUM_TREES = std::atoll(need(i));
        else if (o == "-maxNumGrowCalls") p.MAX_NUM_GROW_CALLS = std::atoll(need(i));
        else if (o == "-formatTest") {                                // prints Double.toString / DecimalFormat renderings
            const double x = std::atof(need(i));
            std::cout << lichee::javaDoubleToString(x) << " " << lichee::decimalFormat(x, 2) << " " << lichee::decimalFormat(x, 3) << "\n";
            return 0;
        }
        else if (o == "-h" || o == "-help") { usage(); return 0; }
        else { std::cerr << "Unrecognized option: " << argv[i] << "\n"; usage(); return -1; }
    }
    // -cp wins over -maxVAFValid wherever it stands on the command line (LineageEngine.java:348-376 applies it last)
    if (p.CP) { p.VAF_MAX = 1.0; p.MAX_ALLOWED_VAF = 1.0; }
    if (a.inputFileName.empty()) { std::cerr << "Required parameter: input file path [-i]\n"; usage(); return -1; }
    if (a.outputFileName.empty()) a.outputFileName = a.inputFileName + ".trees.txt";
    const bool profile = p.INPUT_FORMAT == lichee::Parameters::SNV_WITH_PROFILE;
    if (!haveN && !profile) { std::cerr << "Required parameter: normal sample id [-n]\n"; usage(); return -1; }
    if (!haveAbsent && !profile) { std::cerr << "Required parameter: -maxVAFAbsent\n"; usage(); return -1; }
    if (!havePresent && !profile) { std::cerr << "Required parameter: -minVAFPresent\n"; usage(); return -1; }
    if (!build) { usage(); return -1; }
    try {
        return lichee::buildLineage(a, p);
    } catch (const std::exception &e) {
        std::cerr << e.what() << "\n";
        return 1;
    }
}
