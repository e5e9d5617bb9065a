"""Golden vectors for the rotated-quad crop, made with cv2 in the build container:

    python tests/golden/make_golden_quad_warp.py        # writes tests/golden/quad_warp.npz

cv2.getPerspectiveTransform + cv2.warpPerspective (INTER_LINEAR | WARP_INVERSE_MAP) on a seeded noise page, for
rotated / skewed / partly outside / tiny quads and both border modes.  The patches are stored flattened.
"""
import os

import cv2
import numpy as np


def quads(rng, n, H, W):
    out = []
    for t in range(n):
        cx, cy = rng.uniform(-10, W + 10), rng.uniform(-10, H + 10)
        ww, hh = rng.uniform(2, 90), rng.uniform(2, 40)
        ang = rng.uniform(-0.6, 0.6) if t % 5 else rng.uniform(-3.1, 3.1)
        c, s = np.cos(ang), np.sin(ang)
        q = np.array([[-ww / 2, -hh / 2], [ww / 2, -hh / 2], [ww / 2, hh / 2], [-ww / 2, hh / 2]]) @ np.array(
            [[c, s], [-s, c]]) + [cx, cy]
        q += rng.uniform(-4, 4, q.shape) if t % 3 else 0.0
        out.append(q.reshape(-1))
    # exact axis-aligned integer rectangle (the {32767,0,0,1} weight entry), and a half-pixel one
    out.append(np.array([10, 20, 60, 20, 60, 40, 10, 40], float))
    out.append(np.array([10.5, 20.5, 60.5, 20.5, 60.5, 40.5, 10.5, 40.5], float))
    # degenerate: a patch that rounds to 0x0 and a sliver that rounds to 40x0: no patch
    out.append(np.array([5, 5, 5.2, 5, 5.2, 5.2, 5, 5.2], float))
    out.append(np.array([30, 50, 70, 50.2, 70, 50.4, 30, 50.3], float))
    return np.array(out, np.float32)


def patch_size(q):
    q = q.astype(np.float64).reshape(4, 2)
    e = lambda a, b: np.sqrt((q[b, 0] - q[a, 0]) ** 2 + (q[b, 1] - q[a, 1]) ** 2)
    return int(np.rint(max(e(0, 1), e(3, 2)))), int(np.rint(max(e(0, 3), e(1, 2))))


def main():
    rng = np.random.default_rng(20240)
    H, W = 240, 320
    page = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    qs = quads(rng, 40, H, W)
    sizes, mats, const, repl = [], [], [], []
    for q in qs:
        w, h = patch_size(q)
        if w < 2 or h < 2:  # no patch by definition (the 4-point system is singular)
            sizes.append((0, 0))
            mats.append(np.zeros((3, 3)))
            continue
        rect = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], np.float32)
        m = cv2.getPerspectiveTransform(rect, q.reshape(4, 2))
        sizes.append((w, h))
        mats.append(m)
        flags = cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP
        const.append(cv2.warpPerspective(page, m, (w, h), flags=flags, borderMode=cv2.BORDER_CONSTANT,
                                         borderValue=(7, 7, 7)).reshape(-1))
        repl.append(cv2.warpPerspective(page, m, (w, h), flags=flags, borderMode=cv2.BORDER_REPLICATE).reshape(-1))
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "quad_warp.npz"), seed=20240, page_hw=(H, W),
                        quads=qs, sizes=np.array(sizes, np.int32), mats=np.array(mats), const=np.concatenate(const),
                        repl=np.concatenate(repl), border_value=7, cv2_version=cv2.__version__)
    print(len(qs), "quads,", sum(len(c) for c in const), "bytes per mode")


if __name__ == "__main__":
    main()
