"""GPU end-to-end test of the C++ host + CLI: SSNV file + -clustersFile -> network -> CUDA search ->
.trees.txt, against the Python host (ctypes) on the network the CLI itself dumped."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "lichee_b200", "lichee_b200_build")


def write_inputs(tmp_path, n_clusters, n_samples, seed):
    """A synthetic clonal tree written in the reference's two input formats (README 'Input File
    Types'): one SSNV line per cluster member, one cluster line per cluster."""
    from lichee_b200.synth import make_instance
    net, prm = make_instance(n_clusters, n_samples, seed, complete=False)
    S = net.S + 1                                    # + normal sample (column 0, all zeros)
    snv, clu = ["#chr\tposition\tdescription\tNormal\t" + "\t".join(f"S{i}" for i in range(net.S))], []
    line = 0
    for v in range(1, net.n):
        row = [0.0] + [float(x) for x in net.aaf[v]]
        profile = "".join("1" if x > 0 else "0" for x in row)
        members = []
        for _ in range(min(net.size[v], 3)):
            line += 1
            members.append(str(line))
            snv.append(f"{1 + line % 22}\t{1000 + line}\tsnv{line}\t" + "\t".join(repr(x) for x in row))
        clu.append(profile + "\t" + "\t".join(repr(x) for x in row) + "\t" + ",".join(members))
    (tmp_path / "in.txt").write_text("\n".join(snv) + "\n")
    (tmp_path / "clusters.txt").write_text("\n".join(clu) + "\n")
    return str(tmp_path / "in.txt"), str(tmp_path / "clusters.txt"), S


def parse_trees(path):
    txt = open(path).read()
    trees = []
    for block in re.split(r"\*\*\*\*Tree \d+\*\*\*\*\n", txt)[1:]:
        edges = [tuple(map(int, m)) for m in re.findall(r"^(\d+) -> (\d+)$", block, flags=re.M)]
        score = float(re.search(r"Error score: (\S+)", block).group(1))
        trees.append((edges, score))
    return trees, txt


@pytest.mark.parametrize("complete", [False, True])
def test_cli_end_to_end_matches_python_host(built, tmp_path, complete):
    import lichee_b200
    from conftest import parse_network_dump as parse_dump
    compared = 0
    for seed in range(3, 12):
        inp, clu, S = write_inputs(tmp_path, 14, 5, seed=seed)
        base = [CLI, "-build", "-i", inp, "-clustersFile", clu, "-n", "0", "-maxVAFAbsent", "0.005", "-minVAFPresent", "0.005",
                "-s", "20"] + (["-c"] if complete else [])
        subprocess.run(base + ["-dumpNetwork", str(tmp_path / "net.txt")], check=True, capture_output=True)
        tags, sizes, aaf, adj = parse_dump(tmp_path / "net.txt")
        r = subprocess.run(base + ["-o", str(tmp_path / "out.trees.txt")], check=True, capture_output=True, text=True)
        net = lichee_b200.PHYNetwork(adj, np.array(aaf))
        expect = net.getLineageTrees(numSave=20)
        assert f"Found {net.stats.num_trees} valid tree" in r.stderr
        if net.stats.num_trees == 0:                      # the CLI went on to adjust the network
            assert "after network adjustments" in r.stderr
            continue
        compared += 1
        check_output(tmp_path, expect)
    assert compared >= 2


def check_output(tmp_path, expect):
    got, txt = parse_trees(tmp_path / "out.trees.txt")
    assert len(got) == len(expect) > 0
    for (edges, score), t in zip(got, expect):
        assert score == t.errorScore                       # Double.toString round-trips
        assert sorted(edges) == sorted((p, c) for p, ch in t.treeEdges.items() for c in ch)
    assert re.match(r"Nodes:\n\d+\t[01]+\t\[( [\d.]+)+\]\tsnv\d+", txt) and "Sample decomposition: \n\tSample lineage decomposition: Normal\nGL\n" in txt
    assert "SNV info:\nsnv" in txt


def test_cli_network_adjustment_loop(built, tmp_path):
    """No valid tree -> fixNetwork / ALL_EDGES retry path (LineageEngine.java:110-127) terminates."""
    snv = "#chr\tposition\tdescription\tNormal\tA\tB\n" + "".join(
        f"1\t{i}\ts{i}\t0.0\t{a}\t{b}\n" for i, (a, b) in enumerate([(0.45, 0.45), (0.45, 0.45), (0.44, 0.0), (0.44, 0.0), (0.0, 0.43), (0.0, 0.43)], 1))
    clu = "011\t0.0\t0.2\t0.2\t1,2\n010\t0.0\t0.44\t0.0\t3,4\n001\t0.0\t0.0\t0.43\t5,6\n"
    (tmp_path / "in.txt").write_text(snv)
    (tmp_path / "c.txt").write_text(clu)
    r = subprocess.run([CLI, "-build", "-i", str(tmp_path / "in.txt"), "-clustersFile", str(tmp_path / "c.txt"), "-n", "0",
                        "-maxVAFAbsent", "0.005", "-minVAFPresent", "0.005", "-o", str(tmp_path / "o.txt")],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0
    assert "Found 0 valid tree(s)" in r.stderr and "after network adjustments" in r.stderr


def test_cli_cyclic_constraint_network(built, tmp_path):
    """Three clusters of ONE group whose pairwise centroid differences stay inside -e in one direction only form
    a directed 3-cycle (checkAndAddEdge, PHYNetwork.java:185-232; with -clustersFile no collapse step runs,
    LineageEngine.java:95-99).  The reference's Gabow-Myers search takes any digraph; through the CLI the library
    searches it with the reference-order replay and writes the reference's trees (literal port, oracle/)."""
    import oracle
    from conftest import parse_network_dump as parse_dump
    c = (0.20, 0.20, 0.30)
    a = (c[0] + 0.06, c[1] + 0.06, c[2] - 0.12)
    b = (a[0] - 0.15, a[1] + 0.09, a[2] + 0.06)
    rows = [a, a, b, b, c, c]
    snv = "#chr\tposition\tdescription\tNormal\tA\tB\tC\n" + "".join(
        f"1\t{i}\ts{i}\t0.0\t{r[0]!r}\t{r[1]!r}\t{r[2]!r}\n" for i, r in enumerate(rows, 1))
    clu = "".join(f"0111\t0.0\t{r[0]!r}\t{r[1]!r}\t{r[2]!r}\t{2 * k + 1},{2 * k + 2}\n" for k, r in enumerate((a, b, c)))
    (tmp_path / "in.txt").write_text(snv)
    (tmp_path / "c.txt").write_text(clu)
    base = [CLI, "-build", "-i", str(tmp_path / "in.txt"), "-clustersFile", str(tmp_path / "c.txt"), "-n", "0",
            "-maxVAFAbsent", "0.005", "-minVAFPresent", "0.005", "-e", "0.1", "-s", "10"]
    subprocess.run(base + ["-dumpNetwork", str(tmp_path / "net.txt")], check=True, capture_output=True)
    tags, sizes, aaf, adj = parse_dump(tmp_path / "net.txt")
    inner = {(u, v) for u in range(1, 4) for v in adj[u]}
    assert len(inner) == 3 and all(any((v, w) in inner for w in range(1, 4)) for (_, v) in inner)     # a -> b -> c -> a
    r = subprocess.run(base + ["-o", str(tmp_path / "out.trees.txt")], check=True, capture_output=True, text=True)
    ref = oracle.search(adj, np.array(aaf), 0.1, top_s=10)
    assert ref["n_trees"] > 0 and f"Found {ref['n_trees']} valid tree" in r.stderr
    got, _ = parse_trees(tmp_path / "out.trees.txt")
    assert len(got) == len(ref["scores"])
    for (edges, score), sc, par in zip(got, ref["scores"], ref["parents"]):
        assert score == sc
        assert sorted(edges) == sorted((int(par[v]), v) for v in range(1, len(par)))
