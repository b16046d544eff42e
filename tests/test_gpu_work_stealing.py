"""ONE search shared by two ranks (two processes on the same GPU, gloo): dynamic work stealing with ONE tree cap
(lichee_b200/dist.py).  The merged ranked list and the tree count must equal the single-handle search with the same
cap -- without a binding cap, with the cap inside the first slice, and with the cap in the middle of the order
(PHYNetwork.java:497: one spanningTrees list, one MAX_NUM_TREES)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

CAPS = [10**12, 1000, 150000, 400000]


def _instance():
    from lichee_b200.synth import make_instance
    net, prm = make_instance(12, 5, seed=4)
    return net.adj, np.array(net.aaf), prm.vaf_error_margin


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lichee_b200
    from lichee_b200.dist import run_work_stealing
    store = dist.distributed_c10d._get_default_store()
    adj, vaf, margin = _instance()
    out = []
    for cap in CAPS:
        s = lichee_b200.LineageSearch(adj, vaf, margin, num_save=64, max_num_trees=cap, max_num_grow_calls=1 << 62,
                                      ref_order_budget=-1)
        info = run_work_stealing(s, store, n_slices=24, cap=cap)          # host exchange: both ranks share one GPU
        out.append((cap, info["num_trees"], info["status"], lichee_b200.top_list_sha256(s.trees(), with_seq=False), info["slices_redone"]))
        s.close()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_one_search_one_cap(built):
    import lichee_b200
    adj, vaf, margin = _instance()
    single = {}
    for cap in CAPS:
        s = lichee_b200.LineageSearch(adj, vaf, margin, num_save=64, max_num_trees=cap, max_num_grow_calls=1 << 62,
                                      ref_order_budget=-1)
        st = s.run()
        single[cap] = (int(st.num_trees), int(st.status) & 1, lichee_b200.top_list_sha256(s.trees(), with_seq=False))
        s.close()
    assert single[CAPS[0]][0] > CAPS[-1] > CAPS[-2]                       # the caps really bind
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    redone = 0
    for rank, out in got:
        for cap, num_trees, status, digest, r in out:
            assert (num_trees, status, digest) == single[cap], (rank, cap)
            redone += r
    assert redone > 0                                                     # a binding cap really went through the repair pass
