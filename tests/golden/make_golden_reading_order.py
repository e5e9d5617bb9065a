"""Golden vectors for the reading-order sort, produced by the REAL reference
(detectors/_east/utils.py:500-644).  Run in the build container only:

    python tests/golden/make_golden_reading_order.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def cases():
    rng = np.random.default_rng(123)
    out = []
    for trial in range(24):
        n = int(rng.integers(0, 160)) if trial else 0
        x0 = rng.integers(0, 900, n)
        y0 = rng.integers(0, 500, n)
        w = rng.integers(0, 120, n)
        h = rng.integers(0, 45, n)
        boxes = np.stack([x0, y0, x0 + w, y0 + h], axis=1).astype(np.int64).reshape(-1, 4)
        if trial % 4 == 1 and n > 6:
            boxes[5] = boxes[2]  # duplicates: the reference's dict collapses them
        out.append(boxes)
    # a text-like layout: rows of words with expanded (overlapping) boxes
    rows = []
    for r in range(12):
        x = 20
        for c in range(14):
            w = int(rng.integers(40, 110))
            rows.append((x - 8, 30 + r * 40 - 6 + int(rng.integers(-3, 4)), x + w + 8, 30 + r * 40 + 30))
            x += w + 6
    out.append(np.array(rows, np.int64))
    # larger random pages (the device kernel's grid-based pair generation starts at 256 boxes): duplicates, zero-size
    # boxes, a box over half the page, negative and large coordinates, dense chains of overlaps
    for seed, n, span in [(1, 700, 900), (2, 1500, 2000), (3, 400, 60000), (4, 300, 300)]:
        out.append(large_random(seed, n, span))
    # every box flat: avg_h == 0, the x-gap condition (<= avg_h * inf = NaN) never holds
    r8 = np.random.default_rng(8)
    x0, y0 = r8.integers(0, 500, 300), r8.integers(0, 500, 300)
    out.append(np.stack([x0, y0, x0 + r8.integers(1, 40, 300), y0], axis=1).astype(np.int64))
    return out


def large_random(seed, n, span):
    rng = np.random.default_rng(seed)
    x0 = rng.integers(-50, span, n)
    y0 = rng.integers(-50, span, n)
    w = rng.integers(0, max(8, span // 12), n)
    h = rng.integers(0, max(6, span // 40), n)
    boxes = np.stack([x0, y0, x0 + w, y0 + h], axis=1).astype(np.int64)
    boxes[5] = boxes[17]
    boxes[40] = boxes[17]
    boxes[61, 3] = boxes[61, 1]
    boxes[62] = (0, 0, span // 2, span // 2)
    return boxes


def main():
    ru = refload.utils()
    data = {}
    cs = cases()
    for i, b in enumerate(cs):
        boxes = [tuple(int(v) for v in r) for r in b]
        data[f"boxes_{i}"] = b
        data[f"resolved_{i}"] = np.array(ru.resolve_intersections(boxes), np.int64).reshape(-1, 4)
        data[f"sorted_{i}"] = np.array(ru.sort_boxes_reading_order(boxes), np.int64).reshape(-1, 4)
        data[f"sorted_res_{i}"] = np.array(ru.sort_boxes_reading_order_with_resolutions(boxes), np.int64).reshape(-1, 4)
    data["n_cases"] = np.int64(len(cs))
    np.savez_compressed(os.path.join(OUT, "reading_order.npz"), **data)
    print("wrote reading_order.npz with", len(cs), "cases")


if __name__ == "__main__":
    main()
