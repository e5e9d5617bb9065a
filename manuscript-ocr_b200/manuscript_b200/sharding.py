"""Page sharding across the GPUs of one box.

Pages are independent through decode -> NMS -> crop, so the path shards with no data-path collective:
rank r owns the contiguous page range shard_pages(n_pages, world, r) and results are gathered on the
host.  torch.distributed is only the launcher-provided plumbing (gloo or nccl process group).
"""
from .batch import shard_pages


def gather_page_results(local_results, world_size=None, group=None):
    """local_results: list of per-page picklable results of this rank, in page order of its shard.
    Returns the concatenation over ranks in rank order == global page order (every rank gets it)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return list(local_results)
    world = world_size or dist.get_world_size(group)
    gathered = [None] * world
    dist.all_gather_object(gathered, list(local_results), group=group)
    out = []
    for part in gathered:
        out.extend(part)
    return out


def run_sharded(n_pages, page_fn, group=None):
    """Run page_fn(page_index) for this rank's shard and gather every page's result on the host."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    mine = [page_fn(i) for i in shard_pages(n_pages, world, rank)]
    return gather_page_results(mine, world, group)


__all__ = ["shard_pages", "gather_page_results", "run_sharded"]
