"""Reading order of BASELINE configs[3] pages (10 000 boxes: the global-memory reading-order kernel), ms per 8-page step with
and without the sort:  python profiles/micro/ro_large_bench.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "manuscript-ocr_b200")]
import manuscript_b200 as mb  # noqa: E402
import synthdata  # noqa: E402

dev = torch.device("cuda:0")
s4, g4, i4 = synthdata.make_batch([3, 4], 4096, 10000)
reps8 = [0, 1, 0, 1, 0, 1, 0, 1]
d4 = [torch.from_numpy(np.ascontiguousarray(x[reps8])).to(dev) for x in (s4, g4, i4)]
stream = torch.cuda.current_stream()


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


out = {}
for ro in (0, 1):
    r = mb.PageBatch(device=0, params=mb.EastParams.default(target_size=4096, sort_reading_order=ro), cap_boxes=16384,
                     crops_cap=8 * 12000, out_hw=(32, 128))
    res = r.run(*d4, sync=True)
    out[ro] = timed(lambda: r.run(*d4))
    del r
print(f"RO_LARGE 8 pages x 10000 boxes: {out[0]:.3f} ms without, {out[1]:.3f} ms with the reading-order sort "
      f"(+{out[1] - out[0]:.3f} ms)")
