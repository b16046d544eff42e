"""Multi-GPU plumbing for one sharded search: every rank searches its slice of the canonical
order (LineageSearch(part_rank=rank, part_count=world)), then the ranked lists are exchanged with
ONE all_gather (NCCL over NVLink on the GPU box, gloo in the CPU tests) and merged identically on
every rank.  torch.distributed is plumbing only; search and merge run in libLICHeE_b200.so."""
import torch
import torch.distributed as dist


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


def run_work_stealing(search, store, n_slices, key="lichee_b200/next_slice", group=None, device=None):
    """One sharded search with dynamic load balance: the canonical order is cut into `n_slices`
    contiguous slices (several per GPU) and every rank claims the next unclaimed one with an atomic
    add on the torch.distributed store until none is left, keeping ONE ranked list (and threshold)
    on its GPU; then the lists are merged with one all_gather.  `store` is the c10d store of the
    process group (TCPStore / FileStore: `add` is an atomic fetch-and-add across ranks)."""
    search.run_claimed(n_slices, lambda: store.add(key, 1) - 1)
    return allgather_ranked(search, group=group, device=device)
