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


def gather_to_root(local_results, root=0, group=None):
    """As gather_page_results, but only `root` receives the corpus (BASELINE configs[4]: "results are gathered on the
    host"): returns the list in global page order on root, None elsewhere."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return list(local_results)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    gathered = [None] * world if rank == root else None
    dist.gather_object(list(local_results), gathered, dst=root, group=group)
    if rank != root:
        return None
    out = []
    for part in gathered:
        out.extend(part)
    return out


def prefer_numa_node_of_gpu(device_index):
    """Ask the kernel to place this process's future allocations (the pinned host buffers) on the NUMA node the GPU hangs
    off, even when the CPU set cannot be narrowed (a container whose cpuset shows one socket): set_mempolicy(
    MPOL_PREFERRED, node).  Returns the node, or None when it is unknown or the call is refused (best effort)."""
    import ctypes
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(int(device_index))).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238  # x86_64
        if os.uname().machine != "x86_64":
            return None
        rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64))
        return node if rc == 0 else None
    except Exception:
        return None


def run_sharded(n_pages, page_fn, group=None):
    """Run page_fn(page_index) for this rank's shard and gather every page's result on the host."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    mine = [page_fn(i) for i in shard_pages(n_pages, world, rank)]
    return gather_page_results(mine, world, group)


def bind_to_gpu_cpus(device_index):
    """Pin this process to the CPU cores NVML reports as local to GPU `device_index` (same NUMA node / PCIe root), so
    that the pinned host buffers it allocates afterwards are first-touched next to that GPU.  With one process per
    GPU this keeps every rank's H2D stream on its own socket.  Returns the CPU set used, or None if NVML or the
    affinity mask is unavailable (the binding is an optimisation, never a requirement)."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        allowed = os.sched_getaffinity(0)
        words = (max(allowed) // 64) + 1 if allowed else 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, max(words, (os.cpu_count() or 64) // 64 + 1))
        cpus = {w * 64 + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= allowed
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


__all__ = ["shard_pages", "gather_page_results", "gather_to_root", "run_sharded", "bind_to_gpu_cpus",
           "prefer_numa_node_of_gpu"]
