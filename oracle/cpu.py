"""numpy-facing wrappers over the C oracle, with the reference's function signatures.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Each wrapper names the reference function it restates (paths relative to the reference
root, olegiy/manuscript-ocr v0.1.8).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        L = C.CDLL(path)
        L.orc_polygon_area.restype = C.c_double
        L.orc_polygon_area.argtypes = [_f64p, C.c_int]
        L.orc_line_hit.restype = None
        L.orc_line_hit.argtypes = [_f64p, _f64p, _f64p, _f64p, _f64p]
        L.orc_clip_halfplane.restype = C.c_int
        L.orc_clip_halfplane.argtypes = [_f64p, C.c_int, _f64p, _f64p, _f64p]
        L.orc_polygon_intersection.restype = C.c_int
        L.orc_polygon_intersection.argtypes = [_f64p, C.c_int, _f64p, C.c_int, _f64p]
        L.orc_polygon_iou.restype = C.c_double
        L.orc_polygon_iou.argtypes = [_f64p, _f64p]
        L.orc_should_merge.restype = C.c_int
        L.orc_should_merge.argtypes = [_f64p, _f64p, C.c_double]
        L.orc_normalize_polygon.restype = None
        L.orc_normalize_polygon.argtypes = [_f64p, _f64p, _f64p]
        L.orc_standard_nms.restype = C.c_int64
        L.orc_standard_nms.argtypes = [_f64p, _f64p, C.c_int64, C.c_double, _i64p]
        L.orc_lanms.restype = C.c_int64
        L.orc_lanms.argtypes = [_f32p, C.c_int64, C.c_double, _f32p, C.c_int64,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_decode_quads.restype = C.c_int64
        L.orc_decode_quads.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                       _f32p, C.c_int64]
        L.orc_expand_boxes.restype = None
        L.orc_expand_boxes.argtypes = [_f32p, C.c_int64, C.c_double, C.c_double, _f32p]
        L.orc_scale_boxes.restype = None
        L.orc_scale_boxes.argtypes = [_f32p, C.c_int64, C.c_int, C.c_int, C.c_int]
        L.orc_quad_area_f32.restype = C.c_float
        L.orc_quad_area_f32.argtypes = [_f32p]
        L.orc_point_polygon_test.restype = C.c_int
        L.orc_point_polygon_test.argtypes = [_f32p, C.c_int, C.c_float, C.c_float]
        L.orc_remove_contained.restype = C.c_int64
        L.orc_remove_contained.argtypes = [_f32p, C.c_int64, _u8p]
        L.orc_np_sum_f32.restype = C.c_float
        L.orc_np_sum_f32.argtypes = [_f32p, C.c_int64]
        L.orc_remove_area_anomalies.restype = C.c_int64
        L.orc_remove_area_anomalies.argtypes = [_f32p, C.c_int64, C.c_int, C.c_double, C.c_int64, _u8p]
        L.orc_axis_align.restype = None
        L.orc_axis_align.argtypes = [_f32p, C.c_int64]
        L.orc_word_rect.restype = C.c_int
        L.orc_word_rect.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _i32p]
        L.orc_resize_linear_u8c3.restype = None
        L.orc_resize_linear_u8c3.argtypes = [_u8p, C.c_int, C.c_int, C.c_int64, _u8p, C.c_int, C.c_int]
        L.orc_resize_area_u8c3.restype = None
        L.orc_resize_area_u8c3.argtypes = [_u8p, C.c_int, C.c_int, C.c_int64, _u8p, C.c_int, C.c_int]
        L.orc_resize_plan.restype = None
        L.orc_resize_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _i32p]
        L.orc_crop_resize_pad.restype = C.c_int
        L.orc_crop_resize_pad.argtypes = [_u8p, C.c_int, C.c_int, _i32p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _poly(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---- lanms.py:7-130 primitives -------------------------------------------------------------
def polygon_area(poly):
    p = _poly(poly)
    return float(lib().orc_polygon_area(p, p.shape[0]))


def compute_intersection(p1, p2, A, B):
    out = np.empty(2, np.float64)
    lib().orc_line_hit(_poly(p1), _poly(p2), _poly(A), _poly(B), out)
    return out


def clip_polygon(subject, A, B):
    s = _poly(subject)
    out = np.empty((20, 2), np.float64)
    n = lib().orc_clip_halfplane(s, s.shape[0], _poly(A), _poly(B), out)
    return out[:n].copy(), int(n)


def polygon_intersection(poly1, poly2):
    a, b = _poly(poly1), _poly(poly2)
    out = np.empty((20, 2), np.float64)
    n = lib().orc_polygon_intersection(a, a.shape[0], b, b.shape[0], out)
    return out[:n].copy()


def polygon_iou(poly1, poly2):
    return float(lib().orc_polygon_iou(_poly(poly1), _poly(poly2)))


def should_merge(poly1, poly2, iou_threshold):
    return bool(lib().orc_should_merge(_poly(poly1), _poly(poly2), float(iou_threshold)))


def normalize_polygon(ref, poly):
    out = np.empty((4, 2), np.float64)
    lib().orc_normalize_polygon(_poly(ref), _poly(poly), out)
    return out


# ---- lanms.py:133-153 ------------------------------------------------------------------------
def standard_nms(polys, scores, iou_threshold, return_index=False):
    p = np.ascontiguousarray(polys, dtype=np.float64).reshape(-1, 4, 2)
    s = np.ascontiguousarray(scores, dtype=np.float64).reshape(-1)
    keep = np.empty(max(len(s), 1), np.int64)
    k = lib().orc_standard_nms(p, s, len(s), float(iou_threshold), keep)
    keep = keep[:k]
    if return_index:
        return keep
    return p[keep], s[keep]


# ---- lanms.py:156-207 ------------------------------------------------------------------------
def locality_aware_nms(boxes, iou_threshold, debug=False):
    """Stable-tie restatement of locality_aware_nms.  debug=True also returns the merged
    clusters and the kept cluster ids (cluster id = creation order in x0-sorted scan)."""
    if boxes is None or len(boxes) == 0:
        out = np.zeros((0, 9), np.float32)
        if debug:
            return out, dict(cluster_polys=np.zeros((0, 4, 2)), cluster_scores=np.zeros(0),
                             keep_cluster=np.zeros(0, np.int64))
        return out
    b = np.ascontiguousarray(boxes, dtype=np.float32)
    n = b.shape[0]
    out = np.empty((n, 9), np.float32)
    if debug:
        cp = np.empty((n, 4, 2), np.float64)
        cs = np.empty(n, np.float64)
        nc = C.c_int64(0)
        kc = np.empty(n, np.int64)
        m = lib().orc_lanms(b, n, float(iou_threshold), out, n, cp.ctypes.data, cs.ctypes.data,
                            C.addressof(nc), kc.ctypes.data)
        return out[:m].copy(), dict(cluster_polys=cp[:nc.value].copy(), cluster_scores=cs[:nc.value].copy(),
                                    keep_cluster=kc[:m].copy())
    m = lib().orc_lanms(b, n, float(iou_threshold), out, n, None, None, None, None)
    return out[:m].copy()


# ---- utils.py:328-381 --------------------------------------------------------------------------
def decode_quads_from_maps(score_map, geo_planar, score_thresh, scale, quantization=1):
    """geo_planar is the (8,H,W) network layout; the reference receives its (H,W,8) transposed view
    (infer.py:321) -- same memory."""
    s = np.ascontiguousarray(score_map, dtype=np.float32)
    if s.ndim == 3 and s.shape[0] == 1:
        s = s[0]
    g = np.ascontiguousarray(geo_planar, dtype=np.float32)
    H, W = s.shape
    assert g.shape == (8, H, W)
    q = max(int(quantization), 1)
    cap = ((H + q - 1) // q) * ((W + q - 1) // q)
    out = np.empty((max(cap, 1), 9), np.float32)
    n = lib().orc_decode_quads(s, g, H, W, float(score_thresh), float(scale), q, out, cap)
    if n == -2:
        raise IndexError("quantised pixel index outside the map (utils.py:370 geo_map[y, x])")
    assert n >= 0
    return out[:n].copy()


# ---- utils.py:384-422 --------------------------------------------------------------------------
def expand_boxes(quads, expand_w=0.0, expand_h=0.0):
    q = np.ascontiguousarray(quads, dtype=np.float32).reshape(-1, 9)
    if len(q) == 0 or (expand_w == 0 and expand_h == 0):
        return q
    out = np.empty_like(q)
    lib().orc_expand_boxes(q, len(q), float(expand_w), float(expand_h), out)
    return out


# ---- infer.py:134-233 EAST box filters -----------------------------------------------------------
def scale_boxes_to_original(boxes, orig_size, target_size):
    b = np.array(boxes, dtype=np.float32, copy=True).reshape(-1, 9)
    if len(b):
        lib().orc_scale_boxes(b, len(b), int(orig_size[0]), int(orig_size[1]), int(target_size))
    return b


def point_polygon_test(contour, pt):
    c = np.ascontiguousarray(contour, dtype=np.float32).reshape(-1, 2)
    return int(lib().orc_point_polygon_test(c, len(c), float(pt[0]), float(pt[1])))


def remove_fully_contained_boxes(quads):
    q = np.ascontiguousarray(quads, dtype=np.float32).reshape(-1, 9)
    if len(q) <= 1:
        return q
    keep = np.empty(len(q), np.uint8)
    lib().orc_remove_contained(q, len(q), keep)
    return q[keep.astype(bool)]


def remove_area_anomalies(quads, enabled=True, sigma=5.0, min_count=30):
    q = np.ascontiguousarray(quads, dtype=np.float32).reshape(-1, 9)
    if len(q) == 0:
        return q
    keep = np.empty(len(q), np.uint8)
    lib().orc_remove_area_anomalies(q, len(q), int(bool(enabled)), float(sigma), int(min_count), keep)
    return q[keep.astype(bool)]


def convert_to_axis_aligned(quads):
    q = np.array(quads, dtype=np.float32, copy=True).reshape(-1, 9)
    if len(q):
        lib().orc_axis_align(q, len(q))
    return q


def east_postprocess(quads_nms, orig_size, target_size=1280, expand_w=0.9, expand_h=0.9, axis_aligned=True,
                     remove_anomalies=True, sigma=5.0, min_count=30):
    """infer.py:340-356 in order: expand, scale, contained removal, anomaly removal, axis-align."""
    q = expand_boxes(quads_nms, expand_w, expand_h)
    q = scale_boxes_to_original(q, orig_size, target_size)
    q = remove_fully_contained_boxes(q)
    q = remove_area_anomalies(q, remove_anomalies, sigma, min_count)
    return convert_to_axis_aligned(q) if axis_aligned else q


# ---- _pipeline.py:125-137, 204-221 ----------------------------------------------------------------
def word_rects(polys, img_h, img_w, min_text_size=5):
    """polys (n,4,2) or (n,8+) float -> (rects (n,4) int32 [x1,y1,x2,y2), valid (n,) bool)."""
    if len(polys) == 0:
        return np.zeros((0, 4), np.int32), np.zeros(0, bool)
    p = np.ascontiguousarray(polys, dtype=np.float32).reshape(len(polys), -1)[:, :8].copy()
    rects = np.zeros((len(p), 4), np.int32)
    valid = np.zeros(len(p), bool)
    for i in range(len(p)):
        valid[i] = bool(lib().orc_word_rect(p[i], int(img_h), int(img_w), int(min_text_size), rects[i]))
    return rects, valid


# ---- transforms.py:62-120, 185-193 ------------------------------------------------------------------
def resize_plan(h, w, img_h, img_w):
    plan = np.zeros(5, np.int32)
    lib().orc_resize_plan(int(h), int(w), int(img_h), int(img_w), plan)
    return dict(new_w=int(plan[0]), new_h=int(plan[1]), x0=int(plan[2]), y0=int(plan[3]),
                interp={0: "copy", 1: "linear", 2: "area"}[int(plan[4])])


def cv_resize(img, dsize, interp):
    """cv2.resize(img, (dw,dh), interpolation=INTER_LINEAR|INTER_AREA) for u8 HxWx3."""
    a = np.ascontiguousarray(img, dtype=np.uint8)
    dw, dh = dsize
    sh, sw = a.shape[:2]
    if (dh, dw) == (sh, sw):
        return a.copy()
    out = np.empty((dh, dw, 3), np.uint8)
    fn = lib().orc_resize_linear_u8c3 if interp == "linear" else lib().orc_resize_area_u8c3
    fn(a, sh, sw, sw * 3, out, dh, dw)
    return out


def crop_resize_pad(page, rect, img_h, img_w):
    """One word: page[y1:y2, x1:x2] -> ResizeAndPadA canvas (u8 HWC) and normalised CHW f32."""
    pg = np.ascontiguousarray(page, dtype=np.uint8)
    r = np.ascontiguousarray(rect, dtype=np.int32)
    canvas = np.empty((img_h, img_w, 3), np.uint8)
    chw = np.empty((3, img_h, img_w), np.float32)
    rc = lib().orc_crop_resize_pad(pg, pg.shape[0], pg.shape[1], r, int(img_h), int(img_w),
                                   canvas.ctypes.data, chw.ctypes.data)
    if rc != 0:
        raise ValueError(f"bad crop rect {rect}: {rc}")
    return canvas, chw
