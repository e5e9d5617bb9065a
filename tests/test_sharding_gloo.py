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
