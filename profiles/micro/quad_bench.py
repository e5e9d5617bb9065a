"""Rotated-quad crop micro-benchmark (the bench.py variant `rotated_quad_crop` on fewer pages, plus an A/B check):
    python profiles/micro/quad_bench.py [pages]
Times ms_quad_crop_resize_pad on the benchmark's word boxes turned by up to +-0.15 rad, reports how many quads the staged
kernel took, and compares the whole float32 batch with the generic kernel's (MS_B200_QUAD_NO_STAGE=1) bit for bit."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "manuscript-ocr_b200")]
import manuscript_b200 as mb  # noqa: E402
import synthdata  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S, WORDS, OUT_H, OUT_W = 2048, 2000, 32, 128
dev = torch.device("cuda:0")
score, geo, imgs = synthdata.make_batch(list(range(P)), S, WORDS)
d_score, d_geo, d_pages = (torch.from_numpy(x).to(dev) for x in (score, geo, imgs))
runner = mb.PageBatch(device=0, params=mb.EastParams.default(target_size=S), cap_boxes=2304, want_batch=False)
res = runner.run(d_score, d_geo, None, sync=True)
counts = res.box_counts.cpu().numpy()
bx = res.boxes.cpu().numpy()
quads, page_of = [], []
rng = np.random.default_rng(1)
for pg in range(P):
    q = bx[pg, : counts[pg], :8].reshape(-1, 4, 2).astype(np.float64)
    c = q.mean(axis=1, keepdims=True)
    ang = rng.uniform(-0.15, 0.15, len(q))
    rot = np.stack([np.stack([np.cos(ang), -np.sin(ang)], -1), np.stack([np.sin(ang), np.cos(ang)], -1)], -2)
    quads.append((np.einsum("nij,nkj->nki", rot, q - c) + c).reshape(-1, 8))
    page_of.append(np.full(len(q), pg, np.int32))
quads = torch.from_numpy(np.concatenate(quads).astype(np.float32)).to(dev)
page_of = torch.from_numpy(np.concatenate(page_of)).to(dev)
nq = int(quads.shape[0])
stream = torch.cuda.current_stream()


def run(ctx, out, sizes):
    mb._cabi.check(ctx.lib.ms_quad_crop_resize_pad(ctx.handle, d_pages.data_ptr(), P, S, S, quads.data_ptr(), 8,
                                                   page_of.data_ptr(), nq, 5, 1, 0, OUT_H, OUT_W, out.data_ptr(), None,
                                                   sizes.data_ptr(), C.c_void_p(stream.cuda_stream)))


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ctx = mb._cabi.Context(0)
os.environ["MS_B200_QUAD_NO_STAGE"] = "1"
gctx = mb._cabi.Context(0)
del os.environ["MS_B200_QUAD_NO_STAGE"]
a = torch.empty((nq, 3, OUT_H, OUT_W), dtype=torch.float32, device=dev)
b = torch.empty_like(a)
sa = torch.zeros((nq, 2), dtype=torch.int32, device=dev)
sb = torch.zeros_like(sa)
ms_a = timed(lambda: run(ctx, a, sa))
c = (C.c_int32 * 2)()
mb._cabi.check(ctx.lib.ms_quad_crop_last_counts(ctx.handle, c))
ms_b = timed(lambda: run(gctx, b, sb))
sz = sa.cpu().numpy().astype(np.int64)
alg = 3 * int((sz[:, 0] * sz[:, 1]).sum()) + 3 * OUT_H * OUT_W * 4 * nq
same = bool(torch.equal(a, b)) and bool(torch.equal(sa, sb))
print(f"QUAD pages={P} quads={nq} staged={c[0]} generic={c[1]} ms={ms_a:.3f} ({alg / ms_a / 1e6:.0f} GB/s, "
      f"{alg / ms_a / 1e6 / 6557.4:.3f} of HBM peak) generic-only ms={ms_b:.3f} identical={same}")
if not same:
    diff = (a != b).flatten(1).any(1).nonzero().flatten().cpu().numpy()
    print("differing quads:", len(diff), diff[:20], quads[diff[:3]].cpu().numpy(), sz[diff[:3]])
    sys.exit(1)
