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


def gather_boxes_to_root(counts, rows, root=0, group=None):
    """Host gather of a rank's per-page results without pickling: counts (n_pages_local,) int32 and rows (sum(counts), 9)
    float32 (its pages' boxes, concatenated in page order) go to `root` as two flat CPU tensors over `group` (gloo).
    Returns (counts_all, rows_all) in global page order on root -- ranks own contiguous page ranges, so rank order is page
    order -- and None elsewhere."""
    import numpy as np
    import torch
    import torch.distributed as dist

    counts = np.ascontiguousarray(counts, dtype=np.int32).reshape(-1)
    rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, 9)
    if not (dist.is_available() and dist.is_initialized()):
        return counts, rows
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = torch.tensor([len(counts), len(rows)], dtype=torch.int64)
    all_sizes = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    max_p, max_r = int(max(t[0] for t in all_sizes)), int(max(t[1] for t in all_sizes))
    c_pad = torch.zeros(max_p, dtype=torch.int32)
    c_pad[: len(counts)] = torch.from_numpy(counts)
    r_pad = torch.zeros((max_r, 9), dtype=torch.float32)
    r_pad[: len(rows)] = torch.from_numpy(rows)
    c_all = [torch.empty_like(c_pad) for _ in range(world)] if rank == root else None
    r_all = [torch.empty_like(r_pad) for _ in range(world)] if rank == root else None
    dist.gather(c_pad, c_all, dst=root, group=group)
    dist.gather(r_pad, r_all, dst=root, group=group)
    if rank != root:
        return None
    cs = [c_all[r][: int(all_sizes[r][0])].numpy() for r in range(world)]
    rs = [r_all[r][: int(all_sizes[r][1])].numpy() for r in range(world)]
    return np.concatenate(cs), np.concatenate(rs)


def gather_boxes_via_shm(counts, rows, root=0, group=None, tag="msb200"):
    """gather_boxes_to_root for the ranks of ONE box, through POSIX shared memory instead of sockets: every rank writes
    its counts and rows into its own /dev/shm segment, the root maps them and concatenates (memcpy speed; the process
    group only carries the sizes and two barriers).  Falls back to gather_boxes_to_root when /dev/shm is unusable."""
    import os

    import numpy as np
    import torch
    import torch.distributed as dist

    counts = np.ascontiguousarray(counts, dtype=np.int32).reshape(-1)
    rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, 9)
    if not (dist.is_available() and dist.is_initialized()):
        return counts, rows
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    base = f"/dev/shm/{tag}_{os.environ.get('MASTER_PORT', '0')}_{os.getppid() if world > 1 else os.getpid()}"
    ok = os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK)
    # one collective carries both the sizes and whether every rank can write /dev/shm
    sizes = torch.tensor([len(counts), len(rows), 1 if ok else 0], dtype=torch.int64)
    all_sizes = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    if any(int(sz[2]) == 0 for sz in all_sizes):
        return gather_boxes_to_root(counts, rows, root, group)
    path = f"{base}_{rank}"
    nbytes = counts.nbytes + rows.nbytes
    if rank != root and nbytes:
        mm = np.memmap(path, dtype=np.uint8, mode="w+", shape=(nbytes,))
        mm[: counts.nbytes] = counts.view(np.uint8)
        mm[counts.nbytes:] = rows.reshape(-1).view(np.uint8)
        mm.flush()
        del mm
    dist.barrier(group=group)
    out = None
    if rank == root:
        # every rank's segment is copied ONCE, straight into its place in the result (no per-rank temporaries)
        tot_c = sum(int(sz[0]) for sz in all_sizes)
        tot_r = sum(int(sz[1]) for sz in all_sizes)
        out_c, out_r = np.empty(tot_c, np.int32), np.empty((tot_r, 9), np.float32)
        at_c = at_r = 0
        for r in range(world):
            n_c, n_r = int(all_sizes[r][0]), int(all_sizes[r][1])
            if r == root:
                out_c[at_c:at_c + n_c] = counts
                out_r[at_r:at_r + n_r] = rows
            elif n_c or n_r:
                mm = np.memmap(f"{base}_{r}", dtype=np.uint8, mode="r", shape=(n_c * 4 + n_r * 36,))
                out_c[at_c:at_c + n_c].view(np.uint8)[:] = mm[: n_c * 4]
                out_r[at_r:at_r + n_r].reshape(-1).view(np.uint8)[:] = mm[n_c * 4:]
                del mm
            at_c += n_c
            at_r += n_r
        out = (out_c, out_r)
    dist.barrier(group=group)
    if rank != root and nbytes:
        try:
            os.unlink(path)
        except OSError:
            pass
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


__all__ = ["shard_pages", "gather_page_results", "gather_to_root", "gather_boxes_to_root", "gather_boxes_via_shm", "run_sharded", "bind_to_gpu_cpus",
           "prefer_numa_node_of_gpu"]
