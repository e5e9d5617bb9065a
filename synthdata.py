"""Synthetic EAST outputs (score / geometry maps) and page images of the BASELINE.json shapes.

Shared by tests/ and bench.py.  Not part of the product package.  The recipe follows
SURVEY.md 8(d): word quads on a jittered grid, rotation U(-0.05,0.05) rad, score region =
quad shrunk by 0.3 (as the reference's training targets, detectors/_east/dataset.py:180-199),
in-region score U(0.7,0.99), background score U(0,0.3), geometry = vertex offsets
(vx - col, vy - row) in map pixels + N(0, 0.05) noise, kept smooth 2 map-px beyond the score
region (a real network's geometry head is smooth there; noisy background geometry would
manufacture garbage quads through the quantisation quirk).
"""
import math

import numpy as np

CONFIGS = {
    # name: (page side, words)           BASELINE.json configs[1..3]
    "cfg1": (1280, 500),
    "cfg2": (2048, 2000),
    "cfg3": (4096, 10000),
}


def word_quads(rng, page, n_words, cell_aspect=1.25):
    """(n,4,2) float64 page-pixel quads, TL,TR,BR,BL (positive shoelace in image coordinates)."""
    cols = max(1, int(round(math.sqrt(n_words / cell_aspect))))
    rows = int(math.ceil(n_words / cols))
    cw, ch = page / cols, page / rows
    idx = np.arange(n_words)
    gx, gy = idx % cols, idx // cols
    w = rng.uniform(0.78, 0.96, n_words) * cw
    h = rng.uniform(0.56, 0.80, n_words) * ch
    # left edges aligned per grid column (text columns / line starts) with sub-pixel jitter: the
    # reference sorts candidates by x0 (lanms.py:166), so aligned columns interleave candidates of
    # different words and its merge pass removes only a few percent -- the hard case.
    left = (gx + 0.03) * cw + rng.uniform(-0.5, 0.5, n_words)
    cx = left + w / 2
    cy = (gy + 0.5) * ch + rng.uniform(-0.04, 0.04, n_words) * ch
    ang = rng.uniform(-0.05, 0.05, n_words)
    return _rect(cx, cy, w, h, ang), (cx, cy, w, h, ang)


def _rect(cx, cy, w, h, ang):
    ca, sa = np.cos(ang), np.sin(ang)
    ux = np.stack([-w / 2, w / 2, w / 2, -w / 2], axis=1)
    uy = np.stack([-h / 2, -h / 2, h / 2, h / 2], axis=1)
    x = cx[:, None] + ux * ca[:, None] - uy * sa[:, None]
    y = cy[:, None] + ux * sa[:, None] + uy * ca[:, None]
    return np.stack([x, y], axis=2)


def make_maps(seed, page=1280, n_words=500, stride=4, noise=0.05, halo=2.0, bg_hi=0.3, shrink=0.3):
    """Returns score (M,M) f32, geo (8,M,M) f32 [network layout], gt quads (n,4,2) f64."""
    rng = np.random.default_rng(seed)
    M = page // stride
    quads, (cx, cy, w, h, ang) = word_quads(rng, page, n_words)
    score = rng.uniform(0.0, bg_hi, (M, M)).astype(np.float32)
    geo = np.zeros((8, M, M), np.float32)

    # shrunk score region, in map pixels, in each word's local frame
    r = shrink * np.minimum(w, h)
    hw_s = (w / 2 - r) / stride
    hh_s = (h / 2 - r) / stride
    mcx, mcy = cx / stride, cy / stride
    wx = int(math.ceil(np.max(w) / 2 / stride)) + int(halo) + 2
    wy = int(math.ceil(np.max(h) / 2 / stride)) + int(halo) + 2
    oy, ox = np.meshgrid(np.arange(-wy, wy + 1), np.arange(-wx, wx + 1), indexing="ij")
    pr = np.floor(mcy)[:, None, None].astype(np.int64) + oy[None]  # (n,wy,wx) map rows
    pc = np.floor(mcx)[:, None, None].astype(np.int64) + ox[None]
    dx = pc - mcx[:, None, None]
    dy = pr - mcy[:, None, None]
    ca, sa = np.cos(ang)[:, None, None], np.sin(ang)[:, None, None]
    u = dx * ca + dy * sa
    v = -dx * sa + dy * ca
    inb = (pr >= 0) & (pr < M) & (pc >= 0) & (pc < M)
    in_geo = inb & (np.abs(u) <= hw_s[:, None, None] + halo) & (np.abs(v) <= hh_s[:, None, None] + halo)
    in_score = inb & (np.abs(u) <= hw_s[:, None, None]) & (np.abs(v) <= hh_s[:, None, None])

    wi, ry, rx = np.nonzero(in_geo)
    rr, cc = pr[wi, ry, rx], pc[wi, ry, rx]
    vq = quads / stride
    for k in range(4):
        geo[2 * k, rr, cc] = (vq[wi, k, 0] - cc + rng.normal(0, noise, len(wi))).astype(np.float32)
        geo[2 * k + 1, rr, cc] = (vq[wi, k, 1] - rr + rng.normal(0, noise, len(wi))).astype(np.float32)
    wi, ry, rx = np.nonzero(in_score)
    rr, cc = pr[wi, ry, rx], pc[wi, ry, rx]
    score[rr, cc] = rng.uniform(0.7, 0.99, len(wi)).astype(np.float32)
    return score, geo, quads


def make_page_image(seed, page=1280, channels=3):
    """Random-noise RGB page (worst case for resize rounding; content is irrelevant to the path)."""
    rng = np.random.default_rng(seed + 7919)
    return rng.integers(0, 256, (page, page, channels), dtype=np.uint8)


def make_batch(seeds, page, n_words, with_images=True, shrink=0.3):
    scores, geos, imgs = [], [], []
    for s in seeds:
        sc, ge, _ = make_maps(s, page, n_words, shrink=shrink)
        scores.append(sc)
        geos.append(ge)
        if with_images:
            imgs.append(make_page_image(s, page))
    out = (np.stack(scores), np.stack(geos))
    return out + ((np.stack(imgs),) if with_images else (None,))
