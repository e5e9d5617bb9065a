import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "manuscript-ocr_b200")]
import manuscript_b200 as mb, synthdata
P, S = 64, 2048
score, geo, imgs = synthdata.make_batch(list(range(P)), S, 2000)
d = [torch.from_numpy(x).cuda() for x in (score, geo, imgs)]
stream = torch.cuda.current_stream()
def timed(fn, reps=40, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, env in (("graph+split", {}), ("direct+split", {"MS_B200_NO_GRAPHS": "1"}), ("graph nosplit", {"MS_B200_NO_SPLIT": "1"}), ("direct nosplit", {"MS_B200_NO_GRAPHS": "1", "MS_B200_NO_SPLIT": "1"})):
    os.environ.update(env)
    r = mb.PageBatch(device=0, params=mb.EastParams.default(target_size=S), cap_boxes=2304, crops_cap=P * 2564)
    for k in env: del os.environ[k]
    r.run(*d, sync=True)
    print(name, round(timed(lambda: r.run(*d)), 4), "ms")
    r.ctx.stage_timing(True)
    print(name, "with stage timing", round(timed(lambda: r.run(*d)), 4), "ms")
    r.ctx.stage_timing(False)
    del r
