import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g
    g.build()
    return True


def parse_network_dump(path):
    """Parses the text written by `lichee_b200_build -dumpNetwork`."""
    lines = open(path).read().strip().split("\n")
    n, S = map(int, lines[0].split())
    tags, sizes, aaf, adj = [], [], [], []
    for l in lines[1:]:
        left, right = l.split("|")
        f = left.split()
        tags.append(f[0]); sizes.append(int(f[1])); aaf.append([float(x) for x in f[2:]])
        adj.append([int(x) for x in right.split()])
    assert len(tags) == n and all(len(r) == S for r in aaf)
    return tags, sizes, aaf, adj
