"""Generate tests/golden/*.npz by running the REAL reference (olegiy/manuscript-ocr v0.1.8).

Run in the build container only (needs /root/reference + numpy/numba/cv2):

    python tests/golden/make_golden.py

Inputs are regenerated from seeds by synthdata.py at test time; each fixture stores a sha256 of
its inputs so generator drift is detected, plus the reference's outputs.  Reference functions run
(paths relative to the reference root):
  detectors/_east/utils.py:328  decode_quads_from_maps      detectors/_east/lanms.py:156 locality_aware_nms
  detectors/_east/utils.py:384  expand_boxes                detectors/_east/infer.py:134-233 EAST box filters
  _pipeline.py:204              Pipeline._extract_word_image  recognizers/_trba/data/transforms.py:62 ResizeAndPadA
  cv2.pointPolygonTest / cv2.resize (third-party arithmetic the reference calls)
`lanms_ref` is the unpatched reference (numpy default unstable argsort); `lanms_stable` is the
reference with np.argsort forced to kind="stable" -- the documented tie rule.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import synthdata  # noqa: E402
from oracle import refload  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# (name, seed, page, words, orig_hw): orig_hw is the "original image" size EAST scales back to
PAGES = [
    ("page_s0", 0, 512, 80, (700, 900)),
    ("page_s1", 1, 512, 80, (512, 512)),
    ("page_s2", 2, 640, 150, (2000, 1500)),
    ("page_cfg1", 0, 1280, 500, (1280, 1280)),
]
THR, SCALE, Q, IOU = 0.6, 4.0, 2, 0.2


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def east_post(ep, ru, quads, orig_hw, target):
    ep.target_size = target
    e = ru.expand_boxes(quads, 0.9, 0.9)
    s = ep._scale_boxes_to_original(e, orig_hw)
    c = ep._remove_fully_contained_boxes(s)
    a = ep._remove_area_anomalies(c)
    x = ep._convert_to_axis_aligned(a)
    return e, s, c, a, x


def main():
    import cv2

    ru, rl, rs, T = refload.utils(), refload.lanms(), refload.lanms_stable(), refload.transforms()
    ep = refload.EastPost()
    pc = refload.PipelineCrop(5)

    for name, seed, page, words, orig_hw in PAGES:
        score, geo, _ = synthdata.make_maps(seed, page, words)
        quads = ru.decode_quads_from_maps(score, geo.transpose(1, 2, 0), THR, SCALE, Q)
        nms_ref = rl.locality_aware_nms(quads, IOU)
        nms_stable = rs.locality_aware_nms(quads, IOU)
        e, s, c, a, x = east_post(ep, ru, nms_stable, orig_hw, page)
        # Pipeline crop loop on a noise image of the "original" size (detection order; the reading-order
        # permutation is applied by the caller and does not change crop pixels)
        img = synthdata.make_page_image(seed, max(orig_hw))[: orig_hw[0], : orig_hw[1]]
        rects, valid = [], []
        canv32 = []
        tr32 = T.ResizeAndPadA(img_h=32, img_w=128)
        for row in x:
            poly_list = [tuple(map(float, p)) for p in row[:8].reshape(4, 2)]
            poly = np.array(poly_list, dtype=np.int32)
            xmin, ymin = np.min(poly, axis=0)
            xmax, ymax = np.max(poly, axis=0)
            ok = False
            rect = (0, 0, 0, 0)
            if xmax - xmin >= 5 and ymax - ymin >= 5:
                reg = pc._extract_word_image(img, poly)
                if reg is not None and reg.size > 0:
                    ok = True
                    # locate the view inside img to recover the rectangle
                    off = reg.__array_interface__["data"][0] - img.__array_interface__["data"][0]
                    y1, rem = divmod(off, img.strides[0])
                    x1 = rem // img.strides[1]
                    rect = (x1, y1, x1 + reg.shape[1], y1 + reg.shape[0])
                    if len(canv32) < 24:
                        canv32.append(tr32.apply(reg.copy()))
            rects.append(rect)
            valid.append(ok)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            seed=seed, page=page, words=words, orig_hw=np.array(orig_hw),
            input_sha=sha(score, geo), image_sha=sha(img),
            n_candidates=len(quads), quads_sha=sha(quads),
            quads_head=quads[:64], lanms_ref=nms_ref, lanms_stable=nms_stable,
            expanded=e, scaled=s, contained=c, anomalies=a, aligned=x,
            rects=np.array(rects, np.int32).reshape(-1, 4), valid=np.array(valid, bool),
            canvas32=np.array(canv32, np.uint8).reshape(-1, 32, 128, 3),
        )
        print(name, "N", len(quads), "K", len(nms_stable), "ref==stable", np.array_equal(nms_ref, nms_stable),
              "final", len(x), "crops", int(np.sum(valid)))

    # ---- box-filter stress: nested / duplicate / outlier boxes -------------------------------------
    rng = np.random.default_rng(11)
    n = 220
    cx, cy = rng.uniform(50, 1950, n), rng.uniform(50, 1950, n)
    w, h = rng.uniform(20, 90, n), rng.uniform(10, 40, n)
    base = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy - h / 2, cx + w / 2, cy + h / 2, cx - w / 2, cy + h / 2,
                     rng.uniform(0.6, 0.99, n)], axis=1).astype(np.float32)
    base[:, :8] += rng.normal(0, 1.5, (n, 8)).astype(np.float32)
    inner = base[:40].copy()
    ctr = inner[:, :8].reshape(-1, 4, 2).mean(1, keepdims=True)
    inner[:, :8] = (ctr + 0.5 * (inner[:, :8].reshape(-1, 4, 2) - ctr)).reshape(-1, 8)
    dup = base[40:60].copy()
    big = base[60:63].copy()
    ctr = big[:, :8].reshape(-1, 4, 2).mean(1, keepdims=True)
    big[:, :8] = (ctr + 14.0 * (big[:, :8].reshape(-1, 4, 2) - ctr)).reshape(-1, 8) + 5000.0
    boxes = np.concatenate([base, inner, dup, big]).astype(np.float32)
    boxes = boxes[rng.permutation(len(boxes))]
    ep.target_size = 2048
    c = ep._remove_fully_contained_boxes(boxes)
    a = ep._remove_area_anomalies(c)
    x = ep._convert_to_axis_aligned(a)
    np.savez_compressed(os.path.join(OUT, "box_filters.npz"), boxes=boxes, contained=c, anomalies=a, aligned=x)
    print("box_filters", len(boxes), len(c), len(a))

    # ---- cv2.pointPolygonTest ----------------------------------------------------------------------
    cnts, pts, res = [], [], []
    for it in range(3000):
        if it % 3 == 0:
            cn = rng.integers(0, 8, (4, 2)).astype(np.float32)
            p = rng.integers(0, 8, 2).astype(np.float32)
        elif it % 3 == 1:
            cn = (rng.standard_normal((4, 2)) * 10).astype(np.float32)
            p = (rng.standard_normal(2) * 10).astype(np.float32)
        else:
            cn = (rng.integers(0, 6, (4, 2)) * 0.5).astype(np.float32)
            p = cn[rng.integers(0, 4)] * np.float32(0.5) + cn[rng.integers(0, 4)] * np.float32(0.5)
        cnts.append(cn)
        pts.append(p)
        res.append(int(cv2.pointPolygonTest(cn.reshape(-1, 1, 2), (float(p[0]), float(p[1])), False)))
    np.savez_compressed(os.path.join(OUT, "point_polygon.npz"), contours=np.array(cnts), points=np.array(pts),
                        result=np.array(res, np.int8))

    # ---- ResizeAndPadA on random crop sizes (sha of every canvas, 12 canvases in full) -----------------
    img = synthdata.make_page_image(5, 512)
    rects, shas32, shas64, full = [], [], [], []
    tr64 = T.ResizeAndPadA(img_h=64, img_w=256)
    tr32 = T.ResizeAndPadA(img_h=32, img_w=128)
    sizes = [(1, 1), (1, 300), (200, 1), (32, 128), (64, 256), (16, 64), (64, 200), (33, 10), (31, 128), (32, 129),
             (5, 5), (96, 384), (100, 37), (7, 511), (512, 512), (2, 3)]
    while len(sizes) < 260:
        sizes.append((int(rng.integers(1, 160)), int(rng.integers(1, 420))))
    for (h, w) in sizes:
        y1 = int(rng.integers(0, 512 - h + 1))
        x1 = int(rng.integers(0, 512 - w + 1))
        crop = img[y1:y1 + h, x1:x1 + w]
        c32 = tr32.apply(crop.copy())
        c64 = tr64.apply(crop.copy())
        rects.append((x1, y1, x1 + w, y1 + h))
        shas32.append(sha(c32))
        shas64.append(sha(c64))
        if len(full) < 12:
            full.append(c32)
    np.savez_compressed(os.path.join(OUT, "resize_pad.npz"), image_seed=5, image_sha=sha(img),
                        rects=np.array(rects, np.int32), sha32=np.array(shas32), sha64=np.array(shas64),
                        canvas32=np.array(full, np.uint8))
    print("resize_pad", len(rects))


if __name__ == "__main__":
    if not refload.available():
        sys.exit("reference not found: run in the build container")
    main()
