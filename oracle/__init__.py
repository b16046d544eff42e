"""CPU oracle for the LICHeE lineage-tree search hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import
this package.  lichee_b200/ never does."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblichee_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "lichee_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        if os.path.exists(os.path.join(_HERE, "lichee_oracle.c")):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.lichee_oracle_search.restype = ctypes.c_int
    return _lib


def to_csr(adj):
    off = np.zeros(len(adj) + 1, dtype=np.int32)
    for i, a in enumerate(adj):
        off[i + 1] = off[i] + len(a)
    idx = np.array([c for a in adj for c in a], dtype=np.int32)
    if idx.size == 0:
        idx = np.zeros(1, dtype=np.int32)
    return off, idx


def search(adj, vaf, margin, top_s=1, max_trees=100000, max_grow=100000000, mode=0, want_all=0):
    """Run the restated reference search.  Returns dict(scores, parents, orders, dfs_index,
    n_trees, n_grow, n_checks, stopped[, all_scores, all_parents])."""
    lib = _load()
    vaf = np.ascontiguousarray(vaf, dtype=np.float64)
    n, S = vaf.shape
    off, idx = to_csr(adj)
    sc = np.zeros(max(top_s, 1), dtype=np.float64)
    par = np.full((max(top_s, 1), n), -1, dtype=np.int32)
    order = np.full((max(top_s, 1), n), -1, dtype=np.int32)
    dfs = np.zeros(max(top_s, 1), dtype=np.int64)
    stats = np.zeros(4, dtype=np.int64)
    alls = np.zeros(max(want_all, 1), dtype=np.float64)
    allp = np.zeros((max(want_all, 1), n), dtype=np.int32)
    allo = np.zeros((max(want_all, 1), n), dtype=np.int32)
    P = ctypes.c_void_p
    lib.lichee_oracle_search(
        ctypes.c_int(n), ctypes.c_int(S), P(off.ctypes.data), P(idx.ctypes.data), P(vaf.ctypes.data),
        ctypes.c_double(margin), ctypes.c_long(max_trees), ctypes.c_long(max_grow), ctypes.c_int(mode),
        ctypes.c_int(top_s), P(sc.ctypes.data), P(par.ctypes.data), P(order.ctypes.data), P(dfs.ctypes.data),
        P(stats.ctypes.data), P(alls.ctypes.data if want_all else None), P(allp.ctypes.data if want_all else None),
        ctypes.c_long(want_all), P(allo.ctypes.data if want_all else None))
    m = int(min(stats[0], top_s))
    out = dict(scores=sc[:m], parents=par[:m], orders=order[:m], dfs_index=dfs[:m], n_trees=int(stats[0]),
               n_grow=int(stats[1]), n_checks=int(stats[2]), stopped=bool(stats[3]))
    if want_all:
        k = int(min(stats[0], want_all))
        out["all_scores"] = alls[:k]
        out["all_parents"] = allp[:k]
        out["all_orders"] = allo[:k]
    return out


def canonical_search(adj, vaf, margin, top_s=1, max_trees=100000):
    """Sequential statement of the device's canonical enumeration order (see lichee_oracle.c)."""
    lib = _load()
    vaf = np.ascontiguousarray(vaf, dtype=np.float64)
    n, S = vaf.shape
    off, idx = to_csr(adj)
    sc = np.zeros(max(top_s, 1), dtype=np.float64)
    par = np.full((max(top_s, 1), n), -1, dtype=np.int32)
    seq = np.zeros(max(top_s, 1), dtype=np.int64)
    stats = np.zeros(3, dtype=np.int64)
    sigma = np.zeros(max(n - 1, 1), dtype=np.int32)
    P = ctypes.c_void_p
    lib.lichee_canonical_search(
        ctypes.c_int(n), ctypes.c_int(S), P(off.ctypes.data), P(idx.ctypes.data), P(vaf.ctypes.data),
        ctypes.c_double(margin), ctypes.c_long(max_trees), ctypes.c_int(top_s), P(sc.ctypes.data),
        P(par.ctypes.data), P(seq.ctypes.data), P(stats.ctypes.data), P(sigma.ctypes.data))
    m = int(min(stats[0], top_s))
    return dict(scores=sc[:m], parents=par[:m], seq=seq[:m], n_trees=int(stats[0]), n_grow=int(stats[1]),
                n_checks=int(stats[2]), sigma=sigma[:n - 1])


def canonical_eval_tree(vaf, margin, parent):
    """(valid, score) of one tree under the canonical summation order (see lichee_oracle.c)."""
    lib = _load()
    vaf = np.ascontiguousarray(vaf, dtype=np.float64)
    n, S = vaf.shape
    par = np.ascontiguousarray(parent, dtype=np.int32)
    sc = ctypes.c_double(0.0)
    lib.lichee_canonical_eval_tree.restype = ctypes.c_int
    ok = lib.lichee_canonical_eval_tree(ctypes.c_int(n), ctypes.c_int(S), ctypes.c_void_p(vaf.ctypes.data),
                                        ctypes.c_double(margin), ctypes.c_void_p(par.ctypes.data), ctypes.byref(sc))
    return bool(ok), sc.value
