"""The N>1 path on CPU: two gloo ranks shard a page corpus, run a per-page function on their shard and
gather on the host -- the same plumbing bench.py / a multi-GPU deployment uses with one process per GPU
(no data-path collective).  The per-page work here is the CPU oracle, standing in for the GPU."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pages, out_dir):
    for p in (ROOT, os.path.join(ROOT, "manuscript-ocr_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    import synthdata
    from manuscript_b200.sharding import run_sharded, shard_pages
    from oracle import cpu

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    def page_fn(i):
        score, geo, _ = synthdata.make_maps(100 + i, 256, 20)
        q = cpu.decode_quads_from_maps(score, geo, 0.6, 4.0, 2)
        return (i, rank, cpu.locality_aware_nms(q, 0.2))

    res = run_sharded(n_pages, page_fn)
    mine = list(shard_pages(n_pages, world, rank))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"),
            np.array([[i, r, len(b)] for i, r, b in res] + [[-1, rank, len(mine)]]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_page_sharding(tmp_path):
    import torch.multiprocessing as mp

    world, n_pages = 2, 7
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_pages, str(tmp_path)), nprocs=world, join=True)
    a = np.load(tmp_path / "rank0.npy")
    b = np.load(tmp_path / "rank1.npy")
    np.testing.assert_array_equal(a[:-1], b[:-1])          # every rank holds the same gathered list
    assert list(a[:-1, 0]) == list(range(n_pages))         # in global page order
    assert sorted(set(a[:-1, 1])) == [0, 1]                # produced by both ranks
    assert a[-1, 2] + b[-1, 2] == n_pages                  # shards partition the corpus
    assert (a[:-1, 2] > 0).all()


def _gather_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "manuscript-ocr_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    from manuscript_b200.sharding import gather_boxes_to_root, gather_boxes_via_shm, gather_to_root, shard_pages

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = list(shard_pages(11, world, rank))
    rng = np.random.default_rng(5)
    all_counts = rng.integers(0, 6, 11).astype(np.int32)
    all_rows = rng.random((int(all_counts.sum()), 9)).astype(np.float32)
    starts = np.concatenate([[0], np.cumsum(all_counts)])
    counts = all_counts[mine[0]:mine[-1] + 1]
    rows = all_rows[starts[mine[0]]:starts[mine[-1] + 1]]
    for name, fn in (("tensor", gather_boxes_to_root), ("shm", gather_boxes_via_shm)):
        got = fn(counts, rows)
        if rank == 0:
            assert np.array_equal(got[0], all_counts) and np.array_equal(got[1], all_rows), name
        else:
            assert got is None
    objs = gather_to_root([(i, int(all_counts[i])) for i in mine])
    if rank == 0:
        assert objs == [(i, int(all_counts[i])) for i in range(11)]
        np.save(os.path.join(out_dir, "ok.npy"), np.array([1]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_host_gathers_of_a_sharded_corpus(tmp_path):
    """The three host gathers used for BASELINE configs[4] (pickled objects, flat tensors, shared memory), 3 gloo ranks
    with ragged shards: rank 0 gets every page's rows in global page order."""
    import torch.multiprocessing as mp

    mp.spawn(_gather_worker, args=(3, _free_port(), str(tmp_path)), nprocs=3, join=True)
    assert (tmp_path / "ok.npy").exists()
