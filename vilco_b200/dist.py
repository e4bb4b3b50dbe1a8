"""Multi-GPU plumbing of the Moment-Query path: inference shards BY VIDEO with no data-path collective (every clip is
independent; the reference's DistributedSampler does the same split, MQ/libs/datasets/datasets.py:24).  torch.distributed
is used only for the rendezvous, barriers and the max-over-ranks reduction of timings / the gather of result lists."""
import torch


def shard_indices(n_items, rank, world):
    """indices of the clips rank `rank` owns: r, r + world, r + 2*world, ... (same rule as DistributedSampler without
    shuffling / padding)."""
    return list(range(rank, n_items, world))


def max_over_ranks(values, device=None):
    """element-wise max of a list of python floats over all ranks (identity when not initialised)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def gather_results(local_results):
    """host-side gather of the per-rank result lists (list of dicts with CPU tensors) to every rank, re-interleaved
    into the original clip order."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(local_results)
    world = dist.get_world_size()
    buckets = [None] * world
    dist.all_gather_object(buckets, list(local_results))
    out, i = [], 0
    while any(i < len(b) for b in buckets):
        for b in buckets:
            if i < len(b):
                out.append(b[i])
        i += 1
    return out
