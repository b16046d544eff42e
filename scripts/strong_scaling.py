#!/usr/bin/env python
"""Strong scaling of ONE exhaustive lineage-tree search with dynamic work stealing.

  python scripts/strong_scaling.py --clusters 19 --samples 8                      # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
         --master-port 29533 scripts/strong_scaling.py --clusters 19 --samples 8  # 8 GPUs

The canonical order is cut into --slices-per-gpu x N contiguous slices; every rank claims the next
unclaimed slice with an atomic add on the torch.distributed store, keeps one ranked list on its GPU,
and the lists are merged with one NCCL all_gather (lichee_b200/dist.py).  No cap binds, so the merged
top -s list is bit-identical for every N; the script prints a digest of it to prove that.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clusters", type=int, default=19)
    ap.add_argument("--samples", type=int, default=8)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--save", type=int, default=1000)
    ap.add_argument("--slices-per-gpu", type=int, default=16)
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import lichee_b200
    from lichee_b200.dist import run_work_stealing
    from lichee_b200.synth import make_instance

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    store = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        store = dist.distributed_c10d._get_default_store()
    net, prm = make_instance(args.clusters, args.samples, args.seed)
    s = lichee_b200.LineageSearch(net.adj, net.aaf, prm.vaf_error_margin, num_save=args.save, device=local,
                                  max_num_trees=1 << 60, max_num_grow_calls=1 << 62, ref_order_budget=-1)
    n_slices = args.slices_per_gpu * world
    times = []
    for rep in range(args.repeat):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if world > 1:
            run_work_stealing(s, store, n_slices, key=f"lichee_b200/next_slice/{rep}", device=f"cuda:{local}")
        else:
            s.run()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        st = s.stats()
        cnt = torch.tensor([float(st.num_trees), float(st.num_checks)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt)
        times.append(float(dt.item()))
    trees = s.trees()
    digest = hashlib.sha256(b"".join(np.asarray(t.parent, dtype=np.int32).tobytes() + np.float64(t.errorScore).tobytes()
                                     for t in trees)).hexdigest()[:16]
    if rank == 0:
        print(json.dumps({"n_gpus": world, "clusters": args.clusters, "samples": args.samples, "slices": n_slices,
                          "valid_trees": int(cnt[0].item()), "checks": int(cnt[1].item()), "seconds_best": min(times),
                          "seconds_all": times, "checks_per_s": cnt[1].item() / min(times), "best_score": trees[0].errorScore,
                          "top_list_sha256": digest, "scaling": "strong"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
