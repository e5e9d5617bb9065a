"""Timing of the rotated-quad crop (ms_quad_crop_resize_pad) on the benchmark's page shape: 64 x 2048^2 pages,
2000 word quads per page, device-resident."""
import ctypes as C, json, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "manuscript-ocr_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np, torch
import manuscript_b200 as mb

P, S, K = 64, 2048, 2000
rng = np.random.default_rng(0)
pages = torch.randint(0, 256, (P, S, S, 3), dtype=torch.uint8, device="cuda")
rows = np.zeros((P * K, 8), np.float32)
cx, cy = rng.uniform(60, S - 60, P * K), rng.uniform(30, S - 30, P * K)
ww, hh, a = rng.uniform(100, 200, P * K), rng.uniform(28, 44, P * K), rng.uniform(-0.15, 0.15, P * K)  # bench.py's word boxes: ~150 x 35
c, s = np.cos(a), np.sin(a)
for k, (sx, sy) in enumerate([(-1, -1), (1, -1), (1, 1), (-1, 1)]):
    rows[:, 2 * k] = cx + sx * ww / 2 * c - sy * hh / 2 * s
    rows[:, 2 * k + 1] = cy + sx * ww / 2 * s + sy * hh / 2 * c
page_of = np.repeat(np.arange(P, dtype=np.int32), K)
d_rows, d_po = torch.from_numpy(rows).cuda(), torch.from_numpy(page_of).cuda()
n = P * K
batch = torch.empty((n, 3, 32, 128), dtype=torch.float32, device="cuda")
sizes = torch.zeros((n, 2), dtype=torch.int32, device="cuda")
ctx = mb.Context(0)
st = torch.cuda.current_stream().cuda_stream
def run():
    rc = ctx.lib.ms_quad_crop_resize_pad(ctx.handle, pages.data_ptr(), P, S, S, d_rows.data_ptr(), 8, d_po.data_ptr(), n,
                                         5, 1, 0, 32, 128, batch.data_ptr(), None, sizes.data_ptr(), C.c_void_p(st))
    assert rc == 0
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
sz = sizes.cpu().numpy()
src = float((sz[:, 0].astype(np.int64) * sz[:, 1] * 3).sum())
out = n * 3 * 32 * 128 * 4
print(json.dumps({"quad_crop_ms": ms, "quads": n, "valid": int((sz[:, 0] > 0).sum()), "GBps_algorithmic": (src + out) / ms / 1e6,
                  "frac_of_6557": (src + out) / ms / 1e6 / 6557.4}))
