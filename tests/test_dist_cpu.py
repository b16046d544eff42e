"""world_size-2 gloo test of the multi-GPU host path (no GPU): slice bookkeeping + the single
all_gather that exchanges the ranked lists (lichee_b200.dist.allgather_ranked)."""
import os
import struct

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class FakeParams:
    num_save = 4


class FakeSearch:
    """Stands in for LineageSearch: ranked records in host memory, merged with the same key the
    device uses (score, part, seq)."""
    STRIDE = 16

    def __init__(self, rank):
        self.params = FakeParams()
        rng = np.random.RandomState(10 + rank)
        self.rank = rank
        self.records = [(float(np.round(rng.rand(), 1)), int(i), rank, bytes([rank] * self.STRIDE)) for i in range(3 + rank)]
        self.records.sort(key=lambda r: (r[0], r[2], r[1]))
        self.merged = None

    def record_bytes(self):
        return 24 + self.STRIDE

    def export_ranked(self, ptr=None):
        out = bytearray()
        for k in range(self.params.num_save):
            if k < len(self.records):
                s, q, p, b = self.records[k]
                out += struct.pack("<dqii", s, q, p, 1) + b
            else:
                out += struct.pack("<dqii", float("inf"), 0, 0, 0) + bytes([255] * self.STRIDE)
        return np.frombuffer(bytes(out), dtype=np.uint8)

    def merge_ranked(self, records, n_lists, device_ptr=False):
        rb = self.record_bytes()
        raw = records.tobytes()
        recs = []
        for i in range(n_lists * self.params.num_save):
            s, q, p, valid = struct.unpack_from("<dqii", raw, i * rb)
            if valid:
                recs.append((s, q, p, raw[i * rb + 24:(i + 1) * rb]))
        recs.sort(key=lambda r: (r[0], r[2], r[1]))
        self.merged = recs[:self.params.num_save]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lichee_b200.dist import allgather_ranked
    s = FakeSearch(rank)
    gathered = allgather_ranked(s)
    q.put((rank, [(r[0], r[1], r[2]) for r in s.merged], int(gathered.size)))
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_ranked_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    expected = sorted([r for k in range(world) for r in FakeSearch(k).records][:], key=lambda r: (r[0], r[2], r[1]))[:4]
    for rank, merged, size in got:
        assert size == world * 4 * (24 + 16)
        assert merged == [(r[0], r[1], r[2]) for r in expected]


def test_slices_partition_the_frontier():
    """The slice arithmetic of lichee_search.cu (a = w*R/P, b = w*(R+1)/P) tiles [0, w) exactly."""
    for w in (0, 1, 63, 64, 1000, 12345):
        for P in (1, 2, 3, 8):
            cuts = [(w * r // P, w * (r + 1) // P) for r in range(P)]
            assert cuts[0][0] == 0 and cuts[-1][1] == w
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(P - 1))


class FakeClaimSearch(FakeSearch):
    """Slice bookkeeping only: slice k holds TREES[k] valid trees; a run counts them up to the cap in force."""
    TREES = [5, 0, 7, 3, 9, 2, 8, 1, 4, 6, 3, 3, 5]

    def __init__(self, rank):
        super().__init__(rank)
        self.claimed, self.runs, self.cap = [], [], 10**9
        self.params.max_num_trees = 10**9

    def _count(self, t):
        return min(self.TREES[t % len(self.TREES)], self.cap)

    def run_claimed(self, n_slices, claim, log=None):
        while True:
            t = claim()
            if t >= n_slices:
                break
            self.claimed.append(t)
            if log is not None:
                log.append((t, self._count(t), 0))
        return bool(self.claimed)

    def clear(self):
        self.runs = []

    def set_max_num_trees(self, v):
        self.cap = v

    def run_slice(self, t, n_slices, keep_ranked=False):
        self.runs.append((t, self._count(t)))


def _steal_worker(rank, world, port, q, cap):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lichee_b200.dist import run_work_stealing
    store = dist.distributed_c10d._get_default_store()
    out = []
    for rep in range(2):                       # the same key twice: every call must get a fresh counter
        s = FakeClaimSearch(rank)
        s.cap = s.params.max_num_trees = cap
        info = run_work_stealing(s, store, n_slices=37, cap=cap)
        out.append((s.claimed, s.runs, info["num_trees"], info["status"], info["boundary_slice"]))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def _run_steal(cap, port_base):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = port_base + os.getpid() % 2000
    procs = [ctx.Process(target=_steal_worker, args=(r, world, port, q, cap)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return got


def test_work_stealing_claims_every_slice_exactly_once_gloo_world2():
    got = _run_steal(10**9, 31500)
    for rep in range(2):
        claimed = sorted(t for _, out in got for t in out[rep][0])
        assert claimed == list(range(37))
        total = sum(FakeClaimSearch.TREES[t % 13] for t in range(37))
        assert all(out[rep][2] == total and out[rep][3] == 0 and out[rep][4] is None for _, out in got)
        assert all(out[rep][1] == [] for _, out in got)            # no cap binds: nothing is searched twice


def test_one_global_tree_cap_gloo_world2():
    """Parameters.MAX_NUM_TREES is one cap over the whole search: the slice it falls into is re-run with exactly the
    trees it may still contribute, slices after it are dropped, every rank reports the same total."""
    from lichee_b200.dist import global_cap_plan
    cap = 30
    trees = [FakeClaimSearch.TREES[t % 13] for t in range(37)]
    total, boundary, allowed = global_cap_plan([min(c, cap) for c in trees], cap)
    assert (total, boundary, allowed) == (30, 6, 4)                # 5+0+7+3+9+2 = 26; slice 6 holds 8 and may add 4
    got = _run_steal(cap, 33500)
    for rep in range(2):
        for rank, out in got:
            claimed, runs, num_trees, status, b = out[rep]
            assert num_trees == cap and status == 1 and b == boundary
            if any(t > boundary for t in claimed) or (boundary in claimed and min(trees[boundary], cap) > allowed):
                assert [t for t, _ in runs] == sorted(t for t in claimed if t <= boundary)
                assert all(c == (allowed if t == boundary else min(trees[t], cap)) for t, c in runs)
            else:
                assert runs == []
