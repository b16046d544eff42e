"""Multi-GPU plumbing for ONE search shared by several GPUs.

The canonical order is cut into contiguous slices (several per GPU); every rank claims the next unclaimed slice
with an atomic add on the torch.distributed store, keeps ONE ranked list (and threshold) on its GPU, and the lists
are exchanged with ONE all_gather (NCCL over NVLink on the GPU box, gloo in the CPU tests) and merged identically on
every rank.  Parameters.MAX_NUM_TREES is ONE cap over the whole search (PHYNetwork.java:497 -- one spanningTrees
list): the per-slice tree counts are all-reduced, the slice the cap falls into is re-run with exactly the trees it
may still contribute, slices beyond it are dropped, so the merged answer equals the single-GPU answer for every GPU
count, with or without a binding cap.  torch.distributed is plumbing only; search and merge run in
libLICHeE_b200.so."""
import time

import torch
import torch.distributed as dist

STATUS_TREE_CAP = 1

_generation = {}


def allgather_ranked(search, group=None, device=None):
    """Exchange + merge.  `search` needs record_bytes(), params.num_save, export_ranked(ptr|None)
    and merge_ranked(records, n_lists, device_ptr).  With a CUDA `device` the records never leave
    HBM (export to a CUDA tensor, NCCL all_gather, merge from the gathered tensor)."""
    world = dist.get_world_size(group)
    n = search.record_bytes() * search.params.num_save
    if device is not None and torch.device(device).type == "cuda":
        mine = torch.empty(n, dtype=torch.uint8, device=device)
        search.export_ranked(mine.data_ptr())
        out = torch.empty(world * n, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(out, mine, group=group)
        torch.cuda.synchronize(device)
        search.merge_ranked(out.data_ptr(), world, device_ptr=True)
        return out
    mine = torch.from_numpy(search.export_ranked().copy())
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    out = torch.cat(parts).numpy()
    search.merge_ranked(out, world, device_ptr=False)
    return out


def global_cap_plan(counts, cap):
    """counts[k] = valid trees of slice k (each counted up to `cap`).  Returns (total, boundary, allowed): the number
    of trees of the whole search under ONE cap, the slice the cap falls into (None when it does not bind) and the trees
    that slice may contribute.  Slices after `boundary` contribute nothing."""
    prefix = 0
    for k, c in enumerate(counts):
        if prefix + c >= cap:
            return cap, k, cap - prefix
        prefix += c
    return prefix, None, None


def run_work_stealing(search, store, n_slices, key="lichee_b200/next_slice", group=None, device=None, cap=None):
    """One search shared by every rank of `group`, with dynamic load balance and ONE tree cap.

    store : the c10d store of the process group (TCPStore / FileStore: `add` is an atomic fetch-and-add).
    key   : every call uses a fresh counter (key/generation: all ranks call this the same number of times), so the
            fixNetwork loop, several data sets or a retry can share one process group.
    cap   : Parameters.MAX_NUM_TREES over the WHOLE search; default search.params.max_num_trees.

    Returns a dict: valid trees of the whole search, status, and this rank's share (slices, seconds)."""
    gen = _generation.get(key, 0)
    _generation[key] = gen + 1
    counter = f"{key}/{gen}"
    cap = int(search.params.max_num_trees if cap is None else cap)
    world = dist.get_world_size(group)
    cpu = device is None or torch.device(device).type != "cuda"
    red_dev = "cpu" if cpu else device

    t0 = time.perf_counter()
    log = []
    search.run_claimed(n_slices, lambda: store.add(counter, 1) - 1, log=log)
    t_search = time.perf_counter() - t0

    # per-slice tree counts of the whole search (every slice is claimed by exactly one rank)
    counts = torch.full((n_slices,), -1, dtype=torch.int64, device=red_dev)
    for t, c, _ in log:
        counts[t] = c
    dist.all_reduce(counts, op=dist.ReduceOp.MAX, group=group)
    counts = counts.tolist()
    if min(counts) < 0:
        raise RuntimeError(f"work stealing: slices {[k for k, c in enumerate(counts) if c < 0]} of {n_slices} were never "
                           f"searched (counter {counter!r} on the store was not fresh?)")
    total, boundary, allowed = global_cap_plan(counts, cap)
    t1 = time.perf_counter()
    redone = 0
    if boundary is not None:
        # The cap binds.  A list that saw trees beyond the cap (a slice after the boundary, or the tail of the
        # boundary slice) may have let them displace trees that count: search this rank's slices again, the
        # boundary slice with exactly the trees it may contribute, the ones after it not at all.
        mine = sorted(t for t, _, _ in log)
        if any(t > boundary for t in mine) or (boundary in mine and counts[boundary] > allowed):
            keep = [t for t in mine if t <= boundary]
            search.clear()
            first = True
            for t in keep:
                if t == boundary:
                    search.set_max_num_trees(allowed)
                search.run_slice(t, n_slices, keep_ranked=not first)
                first = False
                redone += 1
            search.set_max_num_trees(cap)
    t_repair = time.perf_counter() - t1
    t2 = time.perf_counter()
    allgather_ranked(search, group=group, device=device)
    t_merge = time.perf_counter() - t2
    return {"num_trees": int(total), "status": STATUS_TREE_CAP if boundary is not None else 0,
            "boundary_slice": boundary, "slices_claimed": [t for t, _, _ in log], "slices_redone": redone,
            "search_s": t_search, "repair_s": t_repair, "merge_s": t_merge, "world": world}
