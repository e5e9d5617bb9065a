"""Function-level mirrors of the reference's hot-path seam, numpy in / numpy out.

Same names, argument meaning and error behaviour as the reference functions they replace; the work
happens in the sm_100a kernels behind the *_host entry points of include/manuscript_b200.h.
Reference paths are relative to the reference root (olegiy/manuscript-ocr v0.1.8).
"""
import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import EastParams, check, default_context


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def _ctx(ctx):
    return ctx if ctx is not None else default_context()


# ---- detectors/_east/utils.py:328-381 ---------------------------------------------------------------------
def decode_quads_from_maps(score_map, geo_map, score_thresh, scale, quantization=1, profile=False, ctx=None):
    """score_map (H,W) f32; geo_map (H,W,8) f32 as the reference passes it (the transposed view of the
    network's (8,H,W) output, infer.py:321) or the planar (8,H,W) array itself.  Returns (N,9) f32 rows
    x0,y0..x3,y3,score in (y,x) order of the (quantised) pixels."""
    del profile  # the reference only prints timings with it
    s = np.asarray(score_map)
    if s.ndim == 3 and s.shape[0] == 1:
        s = s[0]
    if s.ndim != 2:
        raise ValueError(f"score_map must be (H,W), got {s.shape}")
    H, W = s.shape
    g = np.asarray(geo_map)
    if g.shape == (H, W, 8):
        g = g.transpose(2, 0, 1)  # back to the network layout; free when it is the reference's view
    elif g.shape != (8, H, W):
        raise ValueError(f"geo_map must be (H,W,8) or (8,H,W) matching the score map, got {g.shape}")
    s = np.ascontiguousarray(s, dtype=np.float32)
    g = np.ascontiguousarray(g, dtype=np.float32)
    q = max(int(quantization), 1)
    cap = ((H + q - 1) // q) * ((W + q - 1) // q)
    out = np.empty((max(cap, 1), 9), np.float32)
    n = C.c_int64(0)
    cx = _ctx(ctx)
    check(cx.lib.ms_decode_quads_host(cx.handle, _ptr(s), _ptr(g), H, W, float(np.float32(score_thresh)),
                                      float(scale), q, _ptr(out), cap, C.byref(n)))
    return out[: n.value].copy()


def decode_rbox_from_maps(score_map, geo_map, score_thresh, scale, quantization=1, ctx=None):
    """RBOX variant of decode_quads_from_maps -- north_star names it, the reference has no such head (QUAD only), so
    this is an extension with no reference behaviour to match.  geo_map (H,W,5) or (5,H,W): distances to the top,
    right, bottom, left edge of the rotated word rectangle (map units) and its angle; returns (N,9) rows TL,TR,BR,BL +
    score with the thresholding, quantisation and (y,x) order of the QUAD decode."""
    s = np.asarray(score_map)
    if s.ndim == 3 and s.shape[0] == 1:
        s = s[0]
    H, W = s.shape
    g = np.asarray(geo_map)
    if g.shape == (H, W, 5):
        g = g.transpose(2, 0, 1)
    elif g.shape != (5, H, W):
        raise ValueError(f"geo_map must be (H,W,5) or (5,H,W) matching the score map, got {g.shape}")
    s = np.ascontiguousarray(s, dtype=np.float32)
    g = np.ascontiguousarray(g, dtype=np.float32)
    q = max(int(quantization), 1)
    cap = ((H + q - 1) // q) * ((W + q - 1) // q)
    out = np.empty((max(cap, 1), 9), np.float32)
    n = C.c_int64(0)
    cx = _ctx(ctx)
    check(cx.lib.ms_decode_rbox_host(cx.handle, _ptr(s), _ptr(g), H, W, float(np.float32(score_thresh)), float(scale), q,
                                     _ptr(out), cap, C.byref(n)))
    return out[: n.value].copy()


# ---- detectors/_east/lanms.py:156-207 -----------------------------------------------------------------------
def locality_aware_nms(boxes, iou_threshold, ctx=None):
    if boxes is None or len(boxes) == 0:
        return np.zeros((0, 9), dtype=np.float32)
    b = np.ascontiguousarray(boxes, dtype=np.float32)
    if b.ndim != 2 or b.shape[1] != 9:
        raise ValueError(f"boxes must be (n,9), got {b.shape}")
    out = np.empty_like(b)
    m = C.c_int64(0)
    cx = _ctx(ctx)
    check(cx.lib.ms_lanms_host(cx.handle, _ptr(b), b.shape[0], float(iou_threshold), _ptr(out), C.byref(m)))
    return out[: m.value].copy()


# ---- detectors/_east/lanms.py:133-153 -----------------------------------------------------------------------
def standard_nms(polys, scores, iou_threshold, return_index=False, ctx=None):
    p = np.ascontiguousarray(polys, dtype=np.float64).reshape(-1, 4, 2)
    s = np.ascontiguousarray(scores, dtype=np.float64).reshape(-1)
    if len(p) != len(s):
        raise ValueError("polys and scores differ in length")
    keep = np.empty(max(len(s), 1), np.int64)
    k = C.c_int64(0)
    cx = _ctx(ctx)
    check(cx.lib.ms_standard_nms_host(cx.handle, _ptr(p), _ptr(s), len(s), float(iou_threshold), _ptr(keep),
                                      C.byref(k)))
    keep = keep[: k.value].copy()
    if return_index:
        return keep
    return p[keep], s[keep]


# ---- detectors/_east/lanms.py:80-96 ----------------------------------------------------------------------------
def polygon_iou(poly1, poly2, ctx=None):
    """Single pair ((4,2),(4,2)) -> float, or batches ((n,4,2),(n,4,2)) -> (n,) f64."""
    a = np.ascontiguousarray(poly1, dtype=np.float64)
    b = np.ascontiguousarray(poly2, dtype=np.float64)
    single = a.ndim == 2
    a = a.reshape(-1, 4, 2)
    b = b.reshape(-1, 4, 2)
    if a.shape != b.shape:
        raise ValueError("poly1 and poly2 differ in shape")
    out = np.empty(len(a), np.float64)
    cx = _ctx(ctx)
    check(cx.lib.ms_polygon_iou_host(cx.handle, _ptr(a), _ptr(b), len(a), _ptr(out)))
    return float(out[0]) if single else out


def should_merge(poly1, poly2, iou_threshold, ctx=None):
    return polygon_iou(poly1, poly2, ctx=ctx) > iou_threshold


# ---- detectors/_east/utils.py:384-422 --------------------------------------------------------------------------
def expand_boxes(quads, expand_w=0.0, expand_h=0.0, ctx=None):
    q = np.ascontiguousarray(quads, dtype=np.float32)
    if len(q) == 0 or (expand_w == 0 and expand_h == 0):
        return q
    q = q.reshape(-1, 9)
    out = np.empty_like(q)
    cx = _ctx(ctx)
    check(cx.lib.ms_expand_boxes_host(cx.handle, _ptr(q), len(q), float(expand_w), float(expand_h), _ptr(out)))
    return out


# ---- detectors/_east/infer.py:134-233, as sequenced by infer.py:340-356 ------------------------------------------
def east_postprocess(quads_nms, orig_size, target_size=1280, expand_w=0.9, expand_h=0.9, axis_aligned=True,
                     remove_anomalies=True, sigma=5.0, min_count=30, ctx=None):
    """expand_boxes -> _scale_boxes_to_original -> _remove_fully_contained_boxes ->
    _remove_area_anomalies -> _convert_to_axis_aligned.  orig_size = (h, w)."""
    q = np.ascontiguousarray(quads_nms, dtype=np.float32).reshape(-1, 9)
    if len(q) == 0:
        return q
    p = EastParams.default(expand_ratio_w=float(expand_w), expand_ratio_h=float(expand_h),
                           target_size=int(target_size), axis_aligned_output=int(bool(axis_aligned)),
                           remove_area_anomalies=int(bool(remove_anomalies)),
                           anomaly_sigma_threshold=float(sigma), anomaly_min_box_count=int(min_count))
    out = np.empty_like(q)
    m = C.c_int64(0)
    cx = _ctx(ctx)
    check(cx.lib.ms_east_boxes_host(cx.handle, _ptr(q), len(q), C.byref(p), int(orig_size[0]), int(orig_size[1]),
                                    _ptr(out), C.byref(m)))
    return out[: m.value].copy()


# ---- _pipeline.py:125-137, 204-221 ---------------------------------------------------------------------------------
def word_rects(polys, img_h, img_w, min_text_size=5, ctx=None):
    """polys (n,4,2) or (n,>=8) float -> rects (n,4) int32 [x1,y1,x2,y2) and valid (n,) bool: the int32
    truncation, min_text_size filter and clamped slice bounds of the Pipeline crop loop."""
    n = len(polys)
    rects = np.zeros((n, 4), np.int32)
    valid = np.zeros(n, np.uint8)
    if n == 0:
        return rects, valid.astype(bool)
    p = np.ascontiguousarray(np.asarray(polys, dtype=np.float32).reshape(n, -1)[:, :8])
    cx = _ctx(ctx)
    check(cx.lib.ms_word_rects_host(cx.handle, _ptr(p), n, int(img_h), int(img_w), int(min_text_size), _ptr(rects),
                                    _ptr(valid)))
    return rects, valid.astype(bool)


# ---- detectors/_east/utils.py:610-644 + _pipeline.py:105-123 -----------------------------------------------------------
def word_reading_order(polys, ctx=None):
    """polys (n,4,2) or (n,>=8) float -> (n,) int32: index of the word at each reading position, exactly what
    Pipeline.predict's sort + re-match loop produces (duplicates resolved the way the reference's dict / first-match
    loops do).  On the device: up to 4096 boxes and 28672 intersecting pairs in one shared-memory kernel, larger pages
    in the global-memory kernel (up to max(65536, 16 n) intersecting pairs); the exact host restatement beyond that."""
    n = len(polys)
    order = np.zeros(n, np.int32)
    if n == 0:
        return order
    p = np.ascontiguousarray(np.asarray(polys, dtype=np.float32).reshape(n, -1)[:, :8])
    cx = _ctx(ctx)
    rc = cx.lib.ms_reading_order_host(cx.handle, _ptr(p), n, _ptr(order)) if n <= (1 << 20) else _cabi.MS_ERR_CAPACITY
    if rc == _cabi.MS_ERR_CAPACITY:
        # beyond the device kernels' capacity (more than max(65536, 16 n) intersecting pairs): the exact host restatement of the
        # reference's own host logic (this step runs on the host in the reference too; it is not a CPU path of a kernel)
        from .reading_order import host_word_order

        return host_word_order(p)
    check(rc)
    return order


# ---- recognizers/_trba/data/transforms.py:85-120,185-193 + recognizers/_trba/__init__.py:264-288,382-390 ------------
def crop_resize_pad(page, rects, img_h=32, img_w=128, want_canvas=False, want_batch=True, ctx=None):
    """page (H,W,3) u8 RGB, rects (n,4) int32 [x1,y1,x2,y2) -> the TRBA input batch (n,3,img_h,img_w) f32
    ((x/255-0.5)/0.5 of the resize-and-pad canvas) and/or the uint8 canvases (n,img_h,img_w,3)."""
    pg = np.ascontiguousarray(page, dtype=np.uint8)
    if pg.ndim != 3 or pg.shape[2] != 3:
        raise ValueError(f"page must be (H,W,3) uint8, got {pg.shape}")
    r = np.ascontiguousarray(rects, dtype=np.int32).reshape(-1, 4)
    n = len(r)
    batch = np.empty((n, 3, img_h, img_w), np.float32) if want_batch else None
    canvas = np.empty((n, img_h, img_w, 3), np.uint8) if want_canvas else None
    if n:
        cx = _ctx(ctx)
        check(cx.lib.ms_crop_resize_pad_host(cx.handle, _ptr(pg), pg.shape[0], pg.shape[1], _ptr(r), n, int(img_h),
                                             int(img_w), _ptr(batch) if want_batch else None,
                                             _ptr(canvas) if want_canvas else None))
    if want_batch and want_canvas:
        return batch, canvas
    return batch if want_batch else canvas


# ---- rectified crops of rotated quads -- an extension, the reference crops bounding rectangles only -----
_BORDER = {"constant": 0, "replicate": 1}


def warp_quad(page, quad, border="constant", border_value=0, ctx=None):
    """The rectified (h,w,3) u8 patch of one quad (x0,y0..x3,y3: top-left, top-right, bottom-right, bottom-left):
    cv2.warpPerspective(page, cv2.getPerspectiveTransform(rect(w,h), quad), (w,h), INTER_LINEAR | WARP_INVERSE_MAP,
    border) with w,h the rounded longer opposite edges.  None when the quad has no patch (a side < 2 pixels,
    non-finite coordinates or a singular 4-point system)."""
    pg = np.ascontiguousarray(page, dtype=np.uint8)
    if pg.ndim != 3 or pg.shape[2] != 3:
        raise ValueError(f"page must be (H,W,3) uint8, got {pg.shape}")
    q = np.ascontiguousarray(np.asarray(quad, dtype=np.float32).reshape(-1)[:8])
    cx = _ctx(ctx)
    w, h = C.c_int(0), C.c_int(0)
    cap = 1 << 16
    while True:
        patch = np.empty(cap, np.uint8)
        rc = cx.lib.ms_warp_quad_host(cx.handle, _ptr(pg), pg.shape[0], pg.shape[1], _ptr(q), _BORDER[border],
                                      int(border_value), _ptr(patch), cap, C.byref(w), C.byref(h))
        need = w.value * h.value * 3
        if rc == -3 and need > cap:  # MS_ERR_CAPACITY: the size is known now
            cap = need
            continue
        check(rc)
        break
    if need == 0:
        return None
    return patch[:need].reshape(h.value, w.value, 3).copy()


def quad_crop_resize_pad(page, quads, img_h=32, img_w=128, min_text_size=5, border="constant", border_value=0,
                         want_canvas=False, want_batch=True, ctx=None):
    """page (H,W,3) u8, quads (n,4,2) or (n,>=8) float -> (batch and/or canvases as crop_resize_pad, valid (n,) bool).
    Row i belongs to quad i; rows of quads without a patch (or smaller than min_text_size on a side) are all padding
    and valid[i] is False."""
    pg = np.ascontiguousarray(page, dtype=np.uint8)
    if pg.ndim != 3 or pg.shape[2] != 3:
        raise ValueError(f"page must be (H,W,3) uint8, got {pg.shape}")
    n = len(quads)
    q = np.ascontiguousarray(np.asarray(quads, dtype=np.float32).reshape(n, -1)[:, :8]) if n else np.zeros((0, 8), np.float32)
    batch = np.empty((n, 3, img_h, img_w), np.float32) if want_batch else None
    canvas = np.empty((n, img_h, img_w, 3), np.uint8) if want_canvas else None
    sizes = np.zeros((n, 2), np.int32)
    if n:
        cx = _ctx(ctx)
        check(cx.lib.ms_quad_crop_resize_pad_host(cx.handle, _ptr(pg), pg.shape[0], pg.shape[1], _ptr(q), n,
                                                  int(min_text_size), _BORDER[border], int(border_value), int(img_h),
                                                  int(img_w), _ptr(batch) if want_batch else None,
                                                  _ptr(canvas) if want_canvas else None, _ptr(sizes)))
    valid = sizes[:, 0] > 0
    if want_batch and want_canvas:
        return batch, canvas, valid
    return (batch if want_batch else canvas), valid


__all__ = [n for n in dir() if not n.startswith("_") and n not in ("C", "np", "EastParams", "check", "default_context")]
_ = _cabi
