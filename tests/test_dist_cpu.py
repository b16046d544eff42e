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
    def __init__(self, rank):
        super().__init__(rank)
        self.claimed = []

    def run_claimed(self, n_slices, claim):
        while True:
            t = claim()
            if t >= n_slices:
                break
            self.claimed.append(t)
        return bool(self.claimed)


def _steal_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lichee_b200.dist import run_work_stealing
    store = dist.distributed_c10d._get_default_store()
    s = FakeClaimSearch(rank)
    run_work_stealing(s, store, n_slices=37)
    q.put((rank, s.claimed))
    dist.barrier()
    dist.destroy_process_group()


def test_work_stealing_claims_every_slice_exactly_once_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_steal_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    claimed = sorted(t for _, c in got for t in c)
    assert claimed == list(range(37))
