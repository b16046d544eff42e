"""CPU tests of the C++ host mirror (lichee_b200/csrc/lichee_host.cpp, `lichee_b200_build`):
SSNV loading, calling, grouping and constraint-network construction against the independent
Python restatement (lichee_b200/network.py) on every bundled dataset, plus the Java text
formatting the .trees.txt writer depends on.  No GPU: -dumpNetwork stops before the search."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "lichee_b200", "lichee_b200_build")
DATA = "/root/reference/LICHeE/data/"

CFG = {
    "ccRCC/RK26.txt": (0.005, 0.005, {}), "ccRCC/RMH008.txt": (0.005, 0.005, {"min_private_cluster_size": 2}),
    "ccRCC/EV003.txt": (0.005, 0.005, {}), "ccRCC/EV005.txt": (0.005, 0.005, {}), "ccRCC/EV006.txt": (0.005, 0.005, {}),
    "ccRCC/EV007.txt": (0.005, 0.005, {}), "ccRCC/RMH002.txt": (0.005, 0.005, {}), "ccRCC/RMH004.txt": (0.01, 0.01, {}),
    "hgsc/case1.txt": (0.005, 0.01, {}), "hgsc/case2.txt": (0.005, 0.01, {}), "hgsc/case3.txt": (0.005, 0.01, {}),
    "hgsc/case4.txt": (0.005, 0.01, {"min_cluster_size": 3}),
    "hgsc/case5.txt": (0.01, 0.04, {"min_cluster_size": 3, "max_collapse_cluster_diff": 0.1}),
    "hgsc/case6.txt": (0.005, 0.01, {"max_collapse_cluster_diff": 0.1}),
}
FLAG = {"min_private_cluster_size": "-minPrivateClusterSize", "min_cluster_size": "-minClusterSize",
        "max_collapse_cluster_diff": "-maxClusterDist"}


from conftest import parse_network_dump as parse_dump


@pytest.mark.skipif(not os.path.isdir(DATA), reason="bundled datasets live in the read-only reference checkout")
@pytest.mark.parametrize("complete", [False, True])
def test_cpp_network_equals_python_restatement(built, tmp_path, complete):
    from lichee_b200.network import Params, build_network_from_file
    for f, (ab, pr, kw) in CFG.items():
        prm = Params(max_vaf_absent=ab, min_vaf_present=pr, all_edges=complete, **kw)
        net, groups, names = build_network_from_file(DATA + f, prm, gap=1e9)
        out = tmp_path / "net.txt"
        cmd = [CLI, "-build", "-i", DATA + f, "-n", "0", "-maxVAFAbsent", str(ab), "-minVAFPresent", str(pr),
               "-singleCluster", "-dumpNetwork", str(out)] + (["-c"] if complete else [])
        for k, v in kw.items():
            cmd += [FLAG[k], str(v)]
        subprocess.run(cmd, check=True, capture_output=True)
        tags, sizes, aaf, adj = parse_dump(out)
        assert tags == ["GL"] + [net.info[i][0] for i in range(1, net.n)], f
        assert sizes == net.size, f
        assert aaf == [[float(x) for x in row] for row in net.aaf], f      # bit-identical centroids
        assert adj == net.adj, f                                            # same edges in the same list order


def test_java_text_formatting(built):
    def fmt(x):
        return subprocess.run([CLI, "-formatTest", repr(x)], check=True, capture_output=True, text=True).stdout.split()
    # Double.toString: decimal between 1e-3 and 1e7, computerized scientific outside; shortest digits
    assert fmt(0.06421797)[0] == "0.06421797"
    assert fmt(0.001)[0] == "0.001" and fmt(0.0001)[0] == "1.0E-4"
    assert fmt(100.0)[0] == "100.0" and fmt(12345678.0)[0] == "1.2345678E7"
    assert fmt(0.0)[0] == "0.0" and fmt(0.1 + 0.2)[0] == "0.30000000000000004"
    # DecimalFormat("#.##") / ("#.###"): HALF_EVEN, no leading zero below 1, "0" when nothing is left
    assert fmt(0.186)[1:] == [".19", ".186"]
    assert fmt(0.5)[1:] == [".5", ".5"]
    assert fmt(0.0)[1:] == ["0", "0"]
    assert fmt(0.125)[1] == ".12" and fmt(0.375)[1] == ".38"
    assert fmt(2.0)[1] == "2" and fmt(0.004)[1] == "0"


def test_cli_rejects_bad_input(built, tmp_path):
    bad = tmp_path / "bad.txt"
    bad.write_text("#chr\tpos\tdesc\tN\tS1\n1\t10\tx\t0.0\n")
    r = subprocess.run([CLI, "-build", "-i", str(bad), "-n", "0", "-maxVAFAbsent", "0.005", "-minVAFPresent", "0.005",
                        "-singleCluster", "-dumpNetwork", str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 1 and "Wrong input file format" in r.stderr
    r = subprocess.run([CLI, "-build", "-i", str(bad)], capture_output=True, text=True)
    assert r.returncode != 0 and "normal sample id" in r.stderr


def test_readme_example_with_sample_profile_and_clusters_file(built, tmp_path):
    """The example input of the reference README ('Input File Types'): SSNV file with a presence
    profile column (-sampleProfile) and the matching 3-line clusters file (-clustersFile)."""
    snv = ("#chr\tposition\tdescription\tprofile\tNormal\tS1\tS2\tS3\tS4\n"
           "1\t184306474\tA/G HMCN1\t01111\t0.0\t0.1\t0.2\t0.25\t0.15\n"
           "1\t18534005\tC/A IGSF21\t01111\t0.0\t0.1\t0.25\t0.2\t0.1\n"
           "1\t110456920\tG/A UBL4B\t01111\t0.0\t0.4\t0.4\t0.45\t0.45\n"
           "10\t26503064\tC/G MYO3A\t01001\t0.0\t0.4\t0.0\t0.0\t0.24\n")
    clu = ("01111\t0.0\t0.1\t0.23\t0.23\t0.13\t1,2\n"
           "01111\t0.0\t0.4\t0.4\t0.45\t0.45\t3\n"
           "01001\t0.0\t0.4\t0.0\t0.0\t0.24\t4\n")
    (tmp_path / "snv.txt").write_text(snv)
    (tmp_path / "clu.txt").write_text(clu)
    out = tmp_path / "net.txt"
    subprocess.run([CLI, "-build", "-i", str(tmp_path / "snv.txt"), "-sampleProfile", "-clustersFile", str(tmp_path / "clu.txt"),
                    "-dumpNetwork", str(out)], check=True, capture_output=True)
    tags, sizes, aaf, adj = parse_dump(out)
    assert sorted(tags[1:]) == ["01001", "01111", "01111"] and sorted(sizes[1:]) == [1, 1, 2]
    by = {(t, s): row for t, s, row in zip(tags, sizes, aaf)}
    assert by[("01111", 2)] == [0.0, 0.1, 0.23, 0.23, 0.13]
    assert by[("01001", 1)] == [0.0, 0.4, 0.0, 0.0, 0.24]
    n = len(tags)
    indeg = [0] * n
    for ch in adj:
        for c in ch:
            indeg[c] += 1
    assert indeg[0] == 0 and all(d >= 1 for d in indeg[1:])       # every cluster node has a parent
    # the larger 4-sample cluster (0.4 ...) can be the parent of the smaller one and of the 2-sample cluster
    big = tags.index("01111") if sizes[tags.index("01111")] == 1 and aaf[tags.index("01111")][1] == 0.4 else \
        [i for i in range(1, n) if tags[i] == "01111" and aaf[i][1] == 0.4][0]
    assert len(adj[big]) >= 1
