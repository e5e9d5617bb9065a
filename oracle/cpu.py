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
        L.orc_polygon_iou_batch.restype = None
        L.orc_polygon_iou_batch.argtypes = [_f64p, _f64p, C.c_longlong, _f64p]
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
        L.orc_decode_rbox.restype = C.c_int64
        L.orc_decode_rbox.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, _f32p, C.c_int64]
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


def polygon_iou_batch(polys1, polys2):
    """(n,4,2) x (n,4,2) -> (n,) f64, lanms.py:80-91 per pair."""
    a = np.ascontiguousarray(polys1, dtype=np.float64).reshape(-1, 8)
    b = np.ascontiguousarray(polys2, dtype=np.float64).reshape(-1, 8)
    assert a.shape == b.shape
    out = np.empty(len(a), np.float64)
    lib().orc_polygon_iou_batch(a, b, len(a), out)
    return out


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


def decode_rbox_from_maps(score_map, geo5_planar, score_thresh, scale, quantization=1):
    """RBOX decode (NOT a reference behaviour, parity unpinned -- see oracle.c): geo5 (5,H,W) = top / right / bottom /
    left distances + angle; rows = rectangle corners TL,TR,BR,BL + score, thresholding / order as decode_quads."""
    s = np.ascontiguousarray(score_map, dtype=np.float32)
    g = np.ascontiguousarray(geo5_planar, dtype=np.float32)
    H, W = s.shape
    assert g.shape == (5, H, W)
    q = max(int(quantization), 1)
    cap = ((H + q - 1) // q) * ((W + q - 1) // q)
    out = np.empty((max(cap, 1), 9), np.float32)
    n = lib().orc_decode_rbox(s, g, H, W, float(score_thresh), float(scale), q, out, cap)
    if n == -2:
        raise IndexError("quantised pixel index outside the map")
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


# ---- rectified crops of rotated quads.  NOT a reference behaviour (the reference crops the axis-aligned
# ---- bounding rectangle, _pipeline.py:204-221; todo.md:1 lists the rotated crop as future work).  Defined here as
# ----   w, h  = round-half-even of the longer of each pair of opposite edges (float64); no patch if a side is < 2 or > 32767
# ----           or the 4-point system is singular
# ----   Minv  = cv2.getPerspectiveTransform([[0,0],[w-1,0],[w-1,h-1],[0,h-1]], quad)
# ----   patch = cv2.warpPerspective(page, Minv, (w, h), INTER_LINEAR | WARP_INVERSE_MAP, borderMode, borderValue)
# ---- and restated from OpenCV's imgproc/imgwarp.cpp (opencv-python pin >=4.5,<5; checked against cv2 4.13 through
# ---- tests/golden/quad_warp.npz): LU with partial pivoting in float64, the 32x32-blocked coordinate evaluation,
# ---- 5 fractional bits, 15-bit bilinear weights including the {32767,0,0,1} entry of the integer position.
def quad_patch_size(quad):
    q = np.asarray(quad, dtype=np.float32).reshape(-1)[:8].astype(np.float64).reshape(4, 2)

    def edge(a, b):
        dx, dy = q[b, 0] - q[a, 0], q[b, 1] - q[a, 1]
        return np.sqrt(dx * dx + dy * dy)

    with np.errstate(invalid="ignore", over="ignore"):
        ew, eh = max(edge(0, 1), edge(3, 2)), max(edge(0, 3), edge(1, 2))
    if not (np.isfinite(ew) and np.isfinite(eh)):
        return 0, 0
    if ew > 32767.0 or eh > 32767.0:  # beyond the 16-bit coordinates of the remap
        return 0, 0
    w, h = int(np.rint(ew)), int(np.rint(eh))
    return (w, h) if w >= 2 and h >= 2 else (0, 0)  # a 1-pixel side makes the 4-point system singular


def perspective_transform(src, dst):
    """cv2.getPerspectiveTransform(src, dst) (imgwarp.cpp getPerspectiveTransform + matrix_decomp.cpp LUImpl)."""
    src = np.asarray(src, np.float32).reshape(4, 2)
    dst = np.asarray(dst, np.float32).reshape(4, 2)
    a = np.zeros((8, 8))
    b = np.zeros(8)
    with np.errstate(over="ignore", invalid="ignore"):
        for i in range(4):
            a[i][0] = a[i + 4][3] = src[i][0]
            a[i][1] = a[i + 4][4] = src[i][1]
            a[i][2] = a[i + 4][5] = 1
            a[i][6] = np.float32(-src[i][0] * dst[i][0])  # float32 products, as Point2f arithmetic gives
            a[i][7] = np.float32(-src[i][1] * dst[i][0])
            a[i + 4][6] = np.float32(-src[i][0] * dst[i][1])
            a[i + 4][7] = np.float32(-src[i][1] * dst[i][1])
            b[i], b[i + 4] = dst[i][0], dst[i][1]
        eps = np.finfo(np.float64).eps * 100
        for i in range(8):
            k = i
            for j in range(i + 1, 8):
                if abs(a[j][i]) > abs(a[k][i]):
                    k = j
            if abs(a[k][i]) < eps:  # singular (cv2 switches to another solver here): the quad has no patch
                return None
            if k != i:
                a[[i, k], i:] = a[[k, i], i:]
                b[[i, k]] = b[[k, i]]
            d = -1 / a[i][i]
            for j in range(i + 1, 8):
                alpha = a[j][i] * d
                a[j, i + 1:] += alpha * a[i, i + 1:]
                b[j] += alpha * b[i]
        for i in range(7, -1, -1):
            s = b[i]
            for k in range(i + 1, 8):
                s -= a[i][k] * b[k]
            b[i] = s / a[i][i]
    return np.append(b, 1.0).reshape(3, 3)


def warp_perspective_inverse(page, minv, w, h, border="constant", border_value=0):
    """cv2.warpPerspective(page, minv, (w,h), INTER_LINEAR|WARP_INVERSE_MAP, BORDER_CONSTANT|BORDER_REPLICATE) for u8x3."""
    img = np.ascontiguousarray(page, dtype=np.uint8)
    H, W = img.shape[:2]
    M = np.asarray(minv, np.float64).reshape(-1)
    bh0 = min(16, h)
    bw0 = min(1024 // bh0, w)
    ys, xs = np.mgrid[0:h, 0:w]
    bx = (xs // bw0) * bw0
    x1 = xs - bx
    with np.errstate(all="ignore"):
        X0 = (M[0] * bx + M[1] * ys) + M[2]
        Y0 = (M[3] * bx + M[4] * ys) + M[5]
        W0 = (M[6] * bx + M[7] * ys) + M[8]
        Wv = W0 + M[6] * x1
        Wv = np.where(Wv != 0, 32.0 / Wv, 0.0)
        # std::max(INT_MIN, std::min(INT_MAX, v)): a NaN stays in the first argument's slot -> INT_MAX, then INT_MAX
        fX = (X0 + M[0] * x1) * Wv
        fY = (Y0 + M[3] * x1) * Wv
        fX = np.where(np.isnan(fX), 2147483647.0, np.maximum(-2147483648.0, np.minimum(2147483647.0, fX)))
        fY = np.where(np.isnan(fY), 2147483647.0, np.maximum(-2147483648.0, np.minimum(2147483647.0, fY)))
    X = np.rint(fX).astype(np.int64)
    Y = np.rint(fY).astype(np.int64)
    sx = np.clip(X >> 5, -32768, 32767)
    sy = np.clip(Y >> 5, -32768, 32767)
    ax, ay = X & 31, Y & 31
    w00, w01 = (32 - ay) * (32 - ax) * 32, (32 - ay) * ax * 32
    w10, w11 = ay * (32 - ax) * 32, ay * ax * 32
    z = (ax == 0) & (ay == 0)
    w00 = np.where(z, 32767, w00)
    w11 = np.where(z, 1, w11)

    def tap(yy, xx):
        v = img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(np.int64)
        if border != "replicate":
            v[~((yy >= 0) & (yy < H) & (xx >= 0) & (xx < W))] = border_value
        return v

    acc = (tap(sy, sx) * w00[..., None] + tap(sy, sx + 1) * w01[..., None] + tap(sy + 1, sx) * w10[..., None]
           + tap(sy + 1, sx + 1) * w11[..., None])
    return np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def warp_quad(page, quad, border="constant", border_value=0):
    """The rectified (h, w, 3) u8 patch of one quad (x0,y0..x3,y3: top-left, top-right, bottom-right, bottom-left)."""
    w, h = quad_patch_size(quad)
    if w == 0:
        return None
    q = np.asarray(quad, dtype=np.float32).reshape(-1)[:8].reshape(4, 2)
    rect = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], np.float32)
    m = perspective_transform(rect, q)
    return None if m is None else warp_perspective_inverse(page, m, w, h, border, border_value)


def quad_crop_resize_pad(page, quad, img_h, img_w, min_text_size=5, border="constant", border_value=0):
    """warp_quad -> ResizeAndPadA canvas + normalised CHW (as crop_resize_pad); (None, None) when the patch is smaller
    than min_text_size on a side (the rule of _pipeline.py:130-133 applied to the patch)."""
    w, h = quad_patch_size(quad)
    if w < min_text_size or h < min_text_size or w == 0:
        return None, None
    patch = warp_quad(page, quad, border, border_value)
    if patch is None:
        return None, None
    return crop_resize_pad(patch, (0, 0, w, h), img_h, img_w)
