"""Host-side wrapper pieces that need no GPU: the lazy ranked-list sequence hands out the same trees, digests and
list behaviour as a plain list of PHYTree objects built from the same ABI arrays."""
import numpy as np

import lichee_b200
from lichee_b200 import PHYTree, RankedTrees, top_list_sha256


def _arrays():
    sc = np.array([0.0, 0.25, 0.25, 1.5])
    seq = np.array([7, 3, 9, 0], dtype=np.int64)
    par = np.array([[-1, 0, 1, 1], [-1, 0, 0, 2], [-1, 2, 0, 0], [-1, 3, 3, 0]], dtype=np.int32)
    order = np.array([[0, 1, 2, 3], [0, 1, 2, 3], [0, 2, 1, 3], [0, 3, 1, 2]], dtype=np.int32)
    return sc, seq, par, order


def test_ranked_trees_behave_like_the_list_they_replace():
    sc, seq, par, order = _arrays()
    lazy = RankedTrees(sc, seq, par, order)
    plain = [PHYTree(float(sc[k]), par, order, int(seq[k]), k) for k in range(len(sc))]
    assert len(lazy) == 4 and lazy != [] and not (lazy == [])
    assert [(t.errorScore, t.seq, t.parent, t.treeNodes, t.treeEdges) for t in lazy] == \
           [(t.errorScore, t.seq, t.parent, t.treeNodes, t.treeEdges) for t in plain]
    assert lazy[1] is lazy[1] and lazy[-1].parent == plain[-1].parent           # one object per entry, negative indices
    assert [t.seq for t in lazy[1:3]] == [3, 9] and lazy == list(lazy)
    for bad in (4, -5):
        try:
            lazy[bad]
            assert False
        except IndexError:
            pass
    assert top_list_sha256(lazy) == top_list_sha256(plain)
    assert top_list_sha256(lazy, with_seq=False) == top_list_sha256(plain, with_seq=False)
    assert str(lazy[3]) == str(plain[3]) == "0 -> 3\n3 -> 1\n3 -> 2\n"
