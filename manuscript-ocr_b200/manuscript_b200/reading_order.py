"""Reading-order sort of word boxes: host logic between NMS and the crop loop.

Mirrors resolve_intersections / sort_boxes_reading_order / sort_boxes_reading_order_with_resolutions
(reference detectors/_east/utils.py:500-644) result for result, including their quirks (the dict
that collapses boxes with equal compressed coordinates, utils.py:639).  The reference's version is
O(n^2) per sweep in pure Python; this restatement is exact but prunes the work:

  * boxes only ever shrink towards their top-left corner (utils.py:531-542), so two boxes that do
    not intersect initially never intersect later -- each sweep visits only the initially
    intersecting pairs, in the reference's (i, j) order;
  * box centres are multiples of 0.5, so the per-line mean of centres is an exact sum / count and can
    be kept as running sums instead of re-averaging every line for every box (utils.py:588-590).

It stays on the host: the sweep is a sequential recurrence over (i, j) in index order.
"""
import numpy as np


def _intersecting_pairs(b):
    """(i, j) with i < j whose boxes intersect (utils.py:516-519), sorted by (i, j)."""
    n = len(b)
    if n < 2:
        return np.zeros((0, 2), np.int64)
    order = np.argsort(b[:, 0], kind="stable")
    x0s = b[order, 0]
    out = []
    # sweep in x: candidates j have x0_j < x1_i; exact test afterwards
    hi = np.searchsorted(x0s, b[order, 2], side="left")
    for a in range(n):
        i = order[a]
        cand = order[a + 1:hi[a]]
        if len(cand) == 0:
            continue
        bi = b[i]
        c = b[cand]
        ok = ~((bi[2] <= c[:, 0]) | (c[:, 2] <= bi[0]) | (bi[3] <= c[:, 1]) | (c[:, 3] <= bi[1]))
        cj = cand[ok]
        if len(cj):
            lo_ = np.minimum(i, cj)
            hi_ = np.maximum(i, cj)
            out.append(np.stack([lo_, hi_], axis=1))
    if not out:
        return np.zeros((0, 2), np.int64)
    p = np.concatenate(out)
    return p[np.lexsort((p[:, 1], p[:, 0]))]


def resolve_intersections(boxes):
    """utils.py:500-547."""
    resolved = [tuple(int(v) for v in bx) for bx in boxes]
    if len(resolved) < 2:
        return list(boxes) if len(resolved) == 0 else resolved
    b = np.array(resolved, dtype=np.int64).reshape(-1, 4)
    # zero / negative sized boxes can satisfy the reference's predicate without overlapping in the usual
    # sense; the sweep below finds candidates by x0 < x1 only, which covers them as well
    pairs = _intersecting_pairs(b).tolist()
    r = [list(t) for t in resolved]
    for _ in range(50):
        changed = False
        for i, j in pairs:
            a, c = r[i], r[j]
            if not (a[2] <= c[0] or c[2] <= a[0] or a[3] <= c[1] or c[3] <= a[1]):
                a[2] = int(a[2] - (a[2] - a[0]) * 0.1)
                a[3] = int(a[3] - (a[3] - a[1]) * 0.1)
                c[2] = int(c[2] - (c[2] - c[0]) * 0.1)
                c[3] = int(c[3] - (c[3] - c[1]) * 0.1)
                changed = True
        if not changed:
            break
    return [tuple(t) for t in r]


def sort_boxes_reading_order(boxes, y_tol_ratio=0.6, x_gap_ratio=np.inf):
    """utils.py:550-607."""
    if not boxes:
        return []
    boxes = [tuple(bx) for bx in boxes]
    avg_h = float(np.mean([bx[3] - bx[1] for bx in boxes]))
    y_tol = avg_h * y_tol_ratio
    with np.errstate(invalid="ignore"):
        x_gap = float(np.float64(avg_h) * np.float64(x_gap_ratio))
    lines = []  # [members, sum_cy, max_x1]
    for bx in sorted(boxes, key=lambda t: (t[1] + t[3]) / 2):
        cy = (bx[1] + bx[3]) / 2
        placed = False
        for ln in lines:
            line_cy = ln[1] / len(ln[0])
            if abs(cy - line_cy) <= y_tol and (bx[0] - ln[2]) <= x_gap:
                ln[0].append(bx)
                ln[1] += cy
                ln[2] = max(ln[2], bx[2])
                placed = True
                break
        if not placed:
            lines.append([[bx], cy, bx[2]])
    lines.sort(key=lambda ln: ln[1] / len(ln[0]))
    out = []
    for ln in lines:
        out.extend(sorted(ln[0], key=lambda t: t[0]))
    return out


def sort_boxes_reading_order_with_resolutions(boxes, y_tol_ratio=0.6, x_gap_ratio=np.inf):
    """utils.py:610-644 (the dict keeps the LAST original of equal compressed boxes, as the reference's does)."""
    boxes = [tuple(bx) for bx in boxes]
    compressed = resolve_intersections(boxes)
    mapping = {c: o for c, o in zip(compressed, boxes)}
    sorted_compressed = sort_boxes_reading_order(compressed, y_tol_ratio=y_tol_ratio, x_gap_ratio=x_gap_ratio)
    return [mapping[bx] for bx in sorted_compressed]


def int_bbox(polygon):
    """_pipeline.py:105-109: np.array(polygon, dtype=np.int32) truncation, then min / max per axis."""
    poly = np.array(polygon, dtype=np.int32)
    x_min, y_min = np.min(poly, axis=0)
    x_max, y_max = np.max(poly, axis=0)
    return (x_min, y_min, x_max, y_max)


def host_word_order(polys8):
    """(n, 8) float polygons -> (n,) int32, order[r] = index of the word at reading position r: _pipeline.py:105-123 on
    the host (int32 truncation, sort with resolutions, first word of equal integer box).  float32 -> int32 truncation
    is done the way np.array(polygon, dtype=np.int32) does it."""
    p = np.asarray(polys8, dtype=np.float64).reshape(len(polys8), -1)[:, :8]
    ip = p.astype(np.int32)
    xs, ys = ip[:, 0::2], ip[:, 1::2]
    keys = list(zip(xs.min(1).tolist(), ys.min(1).tolist(), xs.max(1).tolist(), ys.max(1).tolist()))
    first = {}
    for i, k in enumerate(keys):
        first.setdefault(k, i)
    return np.array([first[tuple(int(v) for v in bx)] for bx in sort_boxes_reading_order_with_resolutions(keys)],
                    dtype=np.int32)


def reorder_words(words, on_device=True):
    """_pipeline.py:105-123 / infer.py:361-385: words re-ordered to follow the sorted boxes; each sorted box
    picks the FIRST word with equal integer bbox (duplicates resolve the way the reference's loop does).
    Ordered by the CUDA kernel (ms_reading_order_host) up to its capacity, by the exact host restatement beyond
    (this step is host logic in the reference too)."""
    if len(words) == 0:
        return []
    if on_device and all(len(w.polygon) == 4 for w in words):
        from . import ops

        polys = np.array([w.polygon for w in words], dtype=np.float32).reshape(len(words), 8)
        return [words[int(i)] for i in ops.word_reading_order(polys)]
    keys = [tuple(int(v) for v in int_bbox(w.polygon)) for w in words]
    first = {}
    for i, k in enumerate(keys):
        first.setdefault(k, i)
    return [words[first[tuple(int(v) for v in bx)]] for bx in sort_boxes_reading_order_with_resolutions(keys)]
