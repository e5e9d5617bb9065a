"""Golden digests for the shapes bench.py measures (BASELINE.json configs[2] and configs[3]).

Run in the build container only (needs /root/reference + numpy/numba/cv2), once -- minutes of CPU:

    python tests/golden/make_golden_large.py

configs[2]  2048x2048, ~2000 words, seed 0: the REAL reference (detectors/_east/lanms.py:156 locality_aware_nms)
            is run twice on the decoded candidates -- unpatched (numpy's default unstable argsort) and with
            kind="stable" (the documented tie rule) -- and the rows the two disagree on are counted.
            sha256 of candidates / kept rows / final boxes / crop rectangles, plus the uint8 canvases of every
            40th crop (from the reference's ResizeAndPadA, recognizers/_trba/data/transforms.py:62-120).
configs[3]  4096x4096, ~10000 words, seed 3: candidates ~75k, so the reference's O(n^2) Python NMS would run for
            tens of minutes; the C oracle (oracle/oracle.c, pinned to the reference on the small pages) is run
            instead and its digests are committed: candidates, kept rows, final boxes, reading order.
Everything is stored in tests/golden/large_pages.npz (a few hundred kB).
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import synthdata  # noqa: E402
from oracle import cpu, refload  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "large_pages.npz")
THR, SCALE, Q, IOU = 0.6, 4.0, 2, 0.2


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    out = {}
    # ---- configs[2]: live reference ---------------------------------------------------------------------------------
    page, words, seed = 2048, 2000, 0
    score, geo, _ = synthdata.make_maps(seed, page, words)
    ru, rl, rs, T = refload.utils(), refload.lanms(), refload.lanms_stable(), refload.transforms()
    t0 = time.time()
    quads = ru.decode_quads_from_maps(score, geo.transpose(1, 2, 0), THR, SCALE, Q)
    print("cfg2 decode", len(quads), f"{time.time() - t0:.1f}s", flush=True)
    t0 = time.time()
    nms_ref = rl.locality_aware_nms(quads, IOU)
    print("cfg2 reference lanms (unpatched)", len(nms_ref), f"{time.time() - t0:.1f}s", flush=True)
    t0 = time.time()
    nms_stable = rs.locality_aware_nms(quads, IOU)
    print("cfg2 reference lanms (stable)", len(nms_stable), f"{time.time() - t0:.1f}s", flush=True)
    a, b = set(map(bytes, nms_ref)), set(map(bytes, nms_stable))
    x0 = quads[:, 0]
    ties = int(len(x0) - len(np.unique(x0)))
    ep = refload.EastPost(target_size=page)
    e = ru.expand_boxes(nms_stable, 0.9, 0.9)
    final = ep._convert_to_axis_aligned(ep._remove_area_anomalies(ep._remove_fully_contained_boxes(
        ep._scale_boxes_to_original(e, (page, page)))))
    img = synthdata.make_page_image(seed, page)
    pc = refload.PipelineCrop(5)
    tr32 = T.ResizeAndPadA(img_h=32, img_w=128)
    rects, canv = [], []
    for row in final:
        poly = np.array([tuple(map(float, p)) for p in row[:8].reshape(4, 2)], dtype=np.int32)
        xmin, ymin = np.min(poly, axis=0)
        xmax, ymax = np.max(poly, axis=0)
        if xmax - xmin < 5 or ymax - ymin < 5:
            continue
        reg = pc._extract_word_image(img, poly)
        if reg is None or reg.size == 0:
            continue
        off = reg.__array_interface__["data"][0] - img.__array_interface__["data"][0]
        y1, rem = divmod(off, img.strides[0])
        x1 = rem // img.strides[1]
        if len(rects) % 40 == 0:
            canv.append(tr32.apply(reg.copy()))
        rects.append((x1, y1, x1 + reg.shape[1], y1 + reg.shape[0]))
    rects = np.array(rects, np.int32).reshape(-1, 4)
    out.update(
        cfg2_seed=seed, cfg2_page=page, cfg2_words=words, cfg2_input_sha=sha(score, geo), cfg2_image_sha=sha(img),
        cfg2_n_candidates=len(quads), cfg2_quads_sha=sha(quads), cfg2_x0_ties=ties,
        cfg2_lanms_ref_sha=sha(nms_ref), cfg2_lanms_stable_sha=sha(nms_stable), cfg2_n_kept=len(nms_stable),
        cfg2_n_kept_ref=len(nms_ref), cfg2_rows_differing=len(a ^ b), cfg2_rows_only_ref=len(a - b),
        cfg2_lanms_ref=nms_ref, cfg2_lanms_stable=nms_stable,
        cfg2_final_sha=sha(final), cfg2_n_final=len(final), cfg2_rects_sha=sha(rects), cfg2_n_rects=len(rects),
        cfg2_canvas_every40=np.array(canv, np.uint8).reshape(-1, 32, 128, 3),
    )
    print("cfg2: x0 ties", ties, "rows differing ref/stable", len(a ^ b), "final", len(final), "rects", len(rects),
          flush=True)
    # the oracle agrees with the reference here too
    oq = cpu.decode_quads_from_maps(score, geo, THR, SCALE, Q)
    assert np.array_equal(oq, quads)
    assert np.array_equal(cpu.locality_aware_nms(oq, IOU), nms_stable)
    assert np.array_equal(cpu.east_postprocess(nms_stable, (page, page), target_size=page), final)

    # ---- configs[3]: C oracle ---------------------------------------------------------------------------------------
    page, words, seed = 4096, 10000, 3
    score, geo, _ = synthdata.make_maps(seed, page, words)
    t0 = time.time()
    quads = cpu.decode_quads_from_maps(score, geo, THR, SCALE, Q)
    nms = cpu.locality_aware_nms(quads, IOU)
    final = cpu.east_postprocess(nms, (page, page), target_size=page)
    print("cfg3 oracle", len(quads), len(nms), len(final), f"{time.time() - t0:.1f}s", flush=True)
    t0 = time.time()
    # reading order of the final boxes by the REAL reference (utils.py:610-644: O(n^2) Python sweeps, minutes) and the
    # first-match word re-matching of _pipeline.py:105-123
    keys = []
    for row in final:
        poly = np.array([tuple(map(float, p)) for p in row[:8].reshape(4, 2)], dtype=np.int32)
        keys.append((int(poly[:, 0].min()), int(poly[:, 1].min()), int(poly[:, 0].max()), int(poly[:, 1].max())))
    sorted_boxes = ru.sort_boxes_reading_order_with_resolutions(keys)
    first = {}
    for i, k in enumerate(keys):
        first.setdefault(k, i)
    order = np.array([first[tuple(int(v) for v in b)] for b in sorted_boxes], np.int32)
    print("cfg3 reading order", len(order), f"{time.time() - t0:.1f}s", flush=True)
    out.update(
        cfg3_seed=seed, cfg3_page=page, cfg3_words=words, cfg3_input_sha=sha(score, geo),
        cfg3_n_candidates=len(quads), cfg3_quads_sha=sha(quads), cfg3_n_kept=len(nms), cfg3_lanms_sha=sha(nms),
        cfg3_n_final=len(final), cfg3_final_sha=sha(final), cfg3_order_sha=sha(order), cfg3_order=order,
    )
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    if not refload.available():
        sys.exit("reference not found: run in the build container")
    main()
