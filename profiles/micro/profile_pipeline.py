"""cProfile of Pipeline.predict (fused route) on one 2048x2048 page with stub networks: where the host time goes."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "manuscript-ocr_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

import manuscript_b200 as mb
import synthdata

S = 2048
score, geo, imgs = synthdata.make_batch([0], S, 2000)
d_score, d_geo = torch.from_numpy(score).cuda(), torch.from_numpy(geo).cuda()


class Net:
    def __call__(self, x):
        return {"score": d_score[0:1, None], "geometry": d_geo[0:1]}


det = mb.EAST(model=Net(), target_size=S)
rec = mb.TRBA(model=lambda b: [("w", 0.5)] * len(b), img_h=32, img_w=128)
pipe = mb.Pipeline(detector=det, recognizer=rec)
for _ in range(3):
    pipe.predict(imgs[0])
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    pipe.predict(imgs[0])
torch.cuda.synchronize()
print("ms per page", 1e3 * (time.perf_counter() - t0) / 10)
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    pipe.predict(imgs[0])
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
