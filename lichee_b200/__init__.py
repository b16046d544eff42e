"""lichee_b200 -- B200-native lineage-tree search for LICHeE's `-build` path.

Python host mirror of the reference interface for this path (names follow
LICHeE/src/lineage): PHYNetwork.getLineageTrees / evaluateLineageTrees, PHYTree.getErrorScore /
toString, Parameters.  All compute goes through the C ABI in include/lichee_b200.h
(libLICHeE_b200.so, hand-written sm_100a kernels).  There is NO CPU fallback: a missing library or
a missing GPU raises LicheeError.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libLICHeE_b200.so")

MAX_NODES = 128
MAX_SAMPLES = 64
MAX_SAVE = 1024
STATUS_TREE_CAP = 1
STATUS_GROW_CAP = 2
STATUS_CANONICAL_ORDER = 4       # ties / cap prefix / child order follow the canonical order, not the reference's
DEFAULT_REF_ORDER_BUDGET = 1 << 22

EXPORTED_SYMBOLS = [
    "lichee_default_params", "lichee_search_create", "lichee_search_run", "lichee_search_run_slice",
    "lichee_search_clear", "lichee_search_get_stats", "lichee_search_set_max_num_trees",
    "lichee_search_get_tree", "lichee_search_get_trees", "lichee_record_bytes", "lichee_search_export_ranked",
    "lichee_search_merge_ranked", "lichee_search_get_order", "lichee_search_destroy",
    "lichee_get_lineage_trees", "lichee_release_cached_memory", "lichee_last_error", "lichee_version", "lichee_device_count",
]


class LicheeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"lichee_b200 error {code}: {msg}")
        self.code = code


class SearchParams(ctypes.Structure):
    _fields_ = [("vaf_error_margin", ctypes.c_double), ("num_save", ctypes.c_int32),
                ("device", ctypes.c_int32), ("max_num_trees", ctypes.c_int64),
                ("max_num_grow_calls", ctypes.c_int64), ("part_rank", ctypes.c_int32),
                ("part_count", ctypes.c_int32), ("level_capacity", ctypes.c_int64),
                ("ref_order_budget", ctypes.c_int64), ("reserved", ctypes.c_int64 * 3)]


class SearchStats(ctypes.Structure):
    _fields_ = [("num_trees", ctypes.c_int64), ("num_grow_calls", ctypes.c_int64),
                ("num_checks", ctypes.c_int64), ("num_scored", ctypes.c_int64),
                ("num_launches", ctypes.c_int64), ("frontier_bytes", ctypes.c_int64),
                ("kernel_ms_total", ctypes.c_double), ("kernel_ms", ctypes.c_double * 5),
                ("kernel_launches", ctypes.c_int64 * 5), ("kernel_bytes", ctypes.c_int64 * 5),
                ("status", ctypes.c_uint32), ("num_saved", ctypes.c_uint32),
                ("num_evaluated", ctypes.c_int64), ("num_bounded_items", ctypes.c_int64), ("num_leaf_items", ctypes.c_int64)]

    KERNELS = ("check", "scan", "emit", "leaf", "select")

    def as_dict(self):
        d = {}
        for f, _ in self._fields_:
            v = getattr(self, f)
            d[f] = list(v) if hasattr(v, "__len__") else v
        return d


_lib = None


def lib():
    """Loads the C-ABI library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise LicheeError(-100, f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                    "(nvcc, sm_100a).  lichee_b200 has no CPU fallback.")
        L = ctypes.CDLL(_SO)
        L.lichee_last_error.restype = ctypes.c_char_p
        L.lichee_version.restype = ctypes.c_char_p
        L.lichee_record_bytes.restype = ctypes.c_int64
        L.lichee_record_bytes.argtypes = [ctypes.c_void_p]
        L.lichee_search_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32, ctypes.c_int32,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.POINTER(SearchParams)]
        L.lichee_search_run.argtypes = [ctypes.c_void_p]
        L.lichee_search_run_slice.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
        L.lichee_search_clear.argtypes = [ctypes.c_void_p]
        L.lichee_search_set_max_num_trees.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        L.lichee_search_get_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(SearchStats)]
        L.lichee_search_get_tree.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p]
        L.lichee_search_get_trees.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.lichee_search_export_ranked.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.lichee_search_merge_ranked.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int]
        L.lichee_search_get_order.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.lichee_search_destroy.argtypes = [ctypes.c_void_p]
        L.lichee_default_params.argtypes = [ctypes.POINTER(SearchParams)]
        _lib = L
    return _lib


def _ck(rc):
    if rc != 0:
        raise LicheeError(rc, lib().lichee_last_error().decode())


def release_cached_memory():
    """Returns the pooled device buffers / streams of destroyed searches to the driver."""
    _ck(lib().lichee_release_cached_memory())


def default_params():
    p = SearchParams()
    lib().lichee_default_params(ctypes.byref(p))
    return p


def _csr(adj):
    off = np.zeros(len(adj) + 1, dtype=np.int32)
    for i, a in enumerate(adj):
        off[i + 1] = off[i] + len(a)
    idx = np.array([c for a in adj for c in a] or [0], dtype=np.int32)
    return off, idx


class Parameters:
    """Subset of lineage.Parameters that reaches the search (Parameters.java)."""
    VAF_ERROR_MARGIN = 0.1
    MAX_NUM_TREES = 100000
    MAX_NUM_GROW_CALLS = 100000000


class PHYTree:
    """Ranked spanning tree returned by the device (PHYTree.java:44-54).  treeNodes / treeEdges are
    rebuilt on first use from the parent and insertion-order arrays the ABI returns."""

    __slots__ = ("errorScore", "seq", "_par2d", "_ord2d", "_k", "_edges")

    def __init__(self, score, parent, order, seq, k=None):
        # parent / order: this tree's rows, or (k given) the whole ranked list's 2-D arrays, sliced on first use
        self.errorScore = score
        self.seq = seq
        self._par2d = parent
        self._ord2d = order
        self._k = k
        self._edges = None

    @property
    def _parent(self):
        return self._par2d if self._k is None else self._par2d[self._k]

    @property
    def _order(self):
        return self._ord2d if self._k is None else self._ord2d[self._k]

    @property
    def parent(self):
        return self._parent.tolist()

    @property
    def treeNodes(self):
        return [x for x in self._order.tolist() if x >= 0]

    @property
    def treeEdges(self):
        if self._edges is None:
            self._edges = {}
            par = self._parent
            for v in self.treeNodes[1:]:
                self._edges.setdefault(int(par[v]), []).append(v)
        return self._edges

    def getErrorScore(self):
        return self.errorScore

    def toString(self):
        # PHYTree.toString (PHYTree.java:176-185); HashMap<PHYNode,...> iterates small Integer-hashed
        # keys in ascending id order.
        return "".join(f"{p} -> {c}\n" for p in sorted(self.treeEdges) for c in self.treeEdges[p])

    __str__ = toString


class RankedTrees:
    """The ranked list as a read-only sequence over the arrays the ABI filled (scores, sequence numbers, parent and
    insertion-order rows): a PHYTree object is made when an entry is first looked at, so handing back a list of a
    thousand trees costs two array allocations and one library call, not a thousand constructor calls."""

    __slots__ = ("_sc", "_seq", "_par", "_ord", "_made")

    def __init__(self, sc, seq, par, order):
        self._sc, self._seq, self._par, self._ord = sc, seq, par, order
        self._made = {}

    def __len__(self):
        return len(self._sc)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[j] for j in range(*k.indices(len(self)))]
        n = len(self)
        if k < 0:
            k += n
        if not 0 <= k < n:
            raise IndexError("ranked tree index out of range")
        t = self._made.get(k)
        if t is None:
            t = self._made[k] = PHYTree(float(self._sc[k]), self._par, self._ord, int(self._seq[k]), k)
        return t

    def __iter__(self):
        return (self[k] for k in range(len(self)))

    def __eq__(self, other):
        try:
            return len(self) == len(other) and all(a is b for a, b in zip(self, other))
        except TypeError:
            return NotImplemented

    def __repr__(self):
        return f"RankedTrees({len(self)} trees)"


def top_list_sha256(trees, with_seq=True):
    """SHA-256 of a ranked list: scores (f64) | seq (i64) | parents (i32), ranked order.  bench.py prints it for
    every step and tests/golden/canonical_digests.json pins it for the benched shapes.  with_seq=False leaves the
    sequence numbers out: in a sliced search they count from the start of the slice, so lists from different GPU
    counts are compared on scores and trees."""
    import hashlib
    h = hashlib.sha256()
    if isinstance(trees, RankedTrees):       # the same bytes, straight from the arrays
        h.update(np.ascontiguousarray(trees._sc, dtype=np.float64).tobytes())
        if with_seq:
            h.update(np.ascontiguousarray(trees._seq, dtype=np.int64).tobytes())
        h.update(np.ascontiguousarray(trees._par, dtype=np.int32).tobytes())
        return h.hexdigest()
    h.update(np.array([t.errorScore for t in trees], dtype=np.float64).tobytes())
    if with_seq:
        h.update(np.array([t.seq for t in trees], dtype=np.int64).tobytes())
    h.update(np.array([t.parent for t in trees], dtype=np.int32).tobytes())
    return h.hexdigest()


class LineageSearch:
    """One constraint network resident on one GPU (opaque lichee_search handle)."""

    def __init__(self, adj, vaf, vaf_error_margin=Parameters.VAF_ERROR_MARGIN, num_save=1,
                 max_num_trees=Parameters.MAX_NUM_TREES, max_num_grow_calls=Parameters.MAX_NUM_GROW_CALLS,
                 device=0, part_rank=0, part_count=1, level_capacity=0, ref_order_budget=0):
        """ref_order_budget: grow() calls the reference-order replay may spend (0 = default 2^22, < 0 = canonical
        order only); see LICHEE_STATUS_CANONICAL_ORDER in include/lichee_b200.h."""
        vaf = np.ascontiguousarray(vaf, dtype=np.float64)
        if vaf.ndim != 2 or len(adj) != vaf.shape[0]:
            raise LicheeError(-1, f"adjacency lists ({len(adj)}) and centroid matrix {vaf.shape} disagree on the node count")
        self.n, self.S = vaf.shape
        off, idx = _csr(adj)
        p = default_params()
        p.vaf_error_margin = vaf_error_margin
        p.num_save = num_save
        p.device = device
        p.max_num_trees = max_num_trees
        p.max_num_grow_calls = max_num_grow_calls
        p.part_rank, p.part_count = part_rank, part_count
        p.level_capacity = level_capacity
        p.ref_order_budget = ref_order_budget
        self.params = p
        self._h = ctypes.c_void_p()
        _ck(lib().lichee_search_create(ctypes.byref(self._h), self.n, self.S, off.ctypes.data, idx.ctypes.data,
                                       vaf.ctypes.data, ctypes.byref(p)))

    def run(self):
        _ck(lib().lichee_search_run(self._h))
        return self.stats()

    def run_slice(self, slice_id, n_slices, keep_ranked=False):
        _ck(lib().lichee_search_run_slice(self._h, slice_id, n_slices, 1 if keep_ranked else 0))
        return self.stats()

    def run_claimed(self, n_slices, claim, log=None):
        """Dynamic work stealing: `claim()` returns the next unclaimed slice id (or >= n_slices when
        none is left); this GPU keeps one ranked list across everything it claims.  `log`, a list, receives
        (slice, valid trees of the slice, status of the slice) per claimed slice."""
        first = True
        trees_before = 0
        while True:
            t = int(claim())
            if t >= n_slices:
                break
            st = self.run_slice(t, n_slices, keep_ranked=not first)
            if log is not None:
                log.append((t, int(st.num_trees) - trees_before, int(st.status)))
            trees_before = int(st.num_trees)
            first = False
        if first:                       # nothing claimed: an empty ranked list that can still be exported
            _ck(lib().lichee_search_clear(self._h))
        return not first

    def clear(self):
        _ck(lib().lichee_search_clear(self._h))

    def set_max_num_trees(self, v):
        _ck(lib().lichee_search_set_max_num_trees(self._h, int(v)))
        self.params.max_num_trees = int(v)

    def stats(self):
        st = SearchStats()
        _ck(lib().lichee_search_get_stats(self._h, ctypes.byref(st)))
        return st

    def sigma(self):
        s = np.zeros(max(self.n - 1, 1), dtype=np.int32)
        _ck(lib().lichee_search_get_order(self._h, s.ctypes.data))
        return s[:self.n - 1]

    def tree(self, k):
        score = ctypes.c_double()
        seq = ctypes.c_int64()
        par = np.zeros(self.n, dtype=np.int32)
        order = np.zeros(self.n, dtype=np.int32)
        _ck(lib().lichee_search_get_tree(self._h, k, ctypes.byref(score), par.ctypes.data, order.ctypes.data,
                                         ctypes.byref(seq)))
        return PHYTree(float(score.value), par, order, int(seq.value))

    def trees(self):
        m = int(self.stats().num_saved)
        if m == 0:
            return []
        sc = np.zeros(m, dtype=np.float64)
        par = np.zeros((m, self.n), dtype=np.int32)
        order = np.zeros((m, self.n), dtype=np.int32)
        seq = np.zeros(m, dtype=np.int64)
        _ck(lib().lichee_search_get_trees(self._h, 0, m, sc.ctypes.data, par.ctypes.data, order.ctypes.data,
                                          seq.ctypes.data))
        return RankedTrees(sc, seq, par, order)

    def record_bytes(self):
        return int(lib().lichee_record_bytes(self._h))

    def export_ranked(self, dst_ptr=None):
        """Ranked list as raw records.  With dst_ptr (a CUDA device pointer) nothing leaves HBM."""
        if dst_ptr is not None:
            _ck(lib().lichee_search_export_ranked(self._h, ctypes.c_void_p(dst_ptr), 1))
            return None
        buf = np.zeros(self.record_bytes() * self.params.num_save, dtype=np.uint8)
        _ck(lib().lichee_search_export_ranked(self._h, buf.ctypes.data, 0))
        return buf

    def merge_ranked(self, records, n_lists, device_ptr=False):
        ptr = records if device_ptr else records.ctypes.data
        _ck(lib().lichee_search_merge_ranked(self._h, ctypes.c_void_p(ptr), n_lists, 1 if device_ptr else 0))

    def close(self):
        if self._h:
            lib().lichee_search_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PHYNetwork:
    """Constraint network as the search sees it (PHYNetwork.java:65-85): numNodes, numSamples,
    `edges` adjacency lists in reference order and the per-node AAF matrix (PHYNode.getAAF)."""

    def __init__(self, edges, aaf, numSamples=None):
        self.edges = [list(a) for a in edges]
        self.aaf = np.ascontiguousarray(aaf, dtype=np.float64)
        self.numNodes = self.aaf.shape[0]
        self.numSamples = self.aaf.shape[1] if numSamples is None else numSamples
        self.spanningTrees = []
        self.stats = None

    def getLineageTrees(self, numSave=1, device=0, **kw):
        """PHYNetwork.getLineageTrees + evaluateLineageTrees on the GPU: returns the `numSave`
        lowest-error valid trees, ranked; `self.stats.num_trees` is the reference's
        spanningTrees.size()."""
        s = LineageSearch(self.edges, self.aaf, kw.pop("vaf_error_margin", Parameters.VAF_ERROR_MARGIN), numSave,
                          kw.pop("max_num_trees", Parameters.MAX_NUM_TREES),
                          kw.pop("max_num_grow_calls", Parameters.MAX_NUM_GROW_CALLS), device, **kw)
        try:
            self.stats = s.run()
            self.spanningTrees = s.trees()
        finally:
            s.close()
        return self.spanningTrees

    def evaluateLineageTrees(self):
        """Ranking already happened on the device (PHYNetwork.java:726-733)."""
        return self.spanningTrees
