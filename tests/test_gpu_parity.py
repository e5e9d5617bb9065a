"""Parity of the CUDA path (through the C ABI, manuscript_b200.ops) against the committed golden
vectors of the real reference and against the C oracle on seeded inputs.  Bit-exact everywhere:
index sets, f64 IoU decisions, f32 coordinates, uint8 crop pixels (tolerance 0)."""
import hashlib
import os

import numpy as np
import pytest

import synthdata
from oracle import cpu

pytestmark = pytest.mark.gpu

PAGES = ["page_s0", "page_s1", "page_s2", "page_cfg1"]


@pytest.fixture(scope="module")
def ops():
    import manuscript_b200 as mb

    assert os.path.exists(mb.library_path()), "libmanuscript_b200.so not built"
    return mb


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def rows_sorted(a):
    a = np.ascontiguousarray(a)
    return a[np.lexsort(a.T[::-1])]


# ---- known-answer cases of the reference's tests/detectors/east/test_lanms.py:18-188 ---------------------------
SQ4 = np.array([[0, 0], [4, 0], [4, 4], [0, 4]], dtype=np.float64)
SQ4B = np.array([[2, 2], [6, 2], [6, 6], [2, 6]], dtype=np.float64)
UNIT = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=np.float64)


def test_kat_iou_and_merge(ops):
    assert np.isclose(ops.polygon_iou(SQ4, SQ4B), 4 / 28, rtol=1e-5)
    assert ops.polygon_iou(UNIT, UNIT) == pytest.approx(1.0)
    assert ops.polygon_iou(UNIT, UNIT + 2) == pytest.approx(0.0)
    assert ops.should_merge(SQ4, SQ4B, 0.1)
    assert not ops.should_merge(SQ4, SQ4B, 0.2)
    assert not ops.should_merge(UNIT, UNIT, 1.0)
    assert ops.should_merge(UNIT, UNIT, 0.999)
    rev = UNIT[::-1].copy()  # orientation quirk, SURVEY 8a-4
    assert ops.polygon_iou(UNIT, rev) == 0.0
    assert ops.polygon_iou(rev, rev) == 0.0


def test_kat_standard_nms(ops):
    polys = [SQ4, SQ4 + 1, SQ4 + 10]
    kp, ks = ops.standard_nms(polys, [0.9, 0.8, 0.7], 0.1)
    assert len(kp) == 2
    np.testing.assert_array_equal(ks, [0.9, 0.7])


def test_kat_lanms(ops):
    boxes = np.array([[0, 0, 4, 0, 4, 4, 0, 4, 0.9], [1, 1, 5, 1, 5, 5, 1, 5, 0.8],
                      [10, 10, 14, 10, 14, 14, 10, 14, 0.7], [11, 11, 15, 11, 15, 15, 11, 15, 0.6]], np.float32)
    out = ops.locality_aware_nms(boxes, 0.1)
    assert out.shape == (2, 9) and out.dtype == np.float32
    np.testing.assert_array_equal(out, cpu.locality_aware_nms(boxes, 0.1))
    assert ops.locality_aware_nms(np.zeros((0, 9), np.float32), 0.5).shape == (0, 9)
    assert ops.locality_aware_nms(None, 0.5).shape == (0, 9)


# ---- random polygon pairs: IoU bit-exact in f64, including irregular quads ----------------------------------------
def test_polygon_iou_random_bitexact(ops):
    rng = np.random.default_rng(5)
    n = 20000
    base = rng.uniform(0, 50, (n, 1, 2))
    a = base + rng.uniform(-8, 8, (n, 4, 2))          # arbitrary quads: concave, self-intersecting, cw
    b = base + rng.uniform(-8, 8, (n, 4, 2))
    rect = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], float)
    a[: n // 2] = base[: n // 2] + rect * rng.uniform(2, 12, (n // 2, 1, 2))
    b[: n // 2] = base[: n // 2] + rng.uniform(-3, 3, (n // 2, 1, 2)) + rect * rng.uniform(2, 12, (n // 2, 1, 2))
    got = ops.polygon_iou(a, b)
    want = np.array([cpu.polygon_iou(x, y) for x, y in zip(a, b)])
    np.testing.assert_array_equal(got, want)
    assert (want > 0.2).sum() > 1000 and (want == 0).sum() > 100


# ---- whole page chains against the real reference's outputs ---------------------------------------------------------
@pytest.mark.parametrize("name", PAGES)
def test_page_chain_golden(ops, golden_dir, name):
    g = load(golden_dir, name)
    seed, page, words = int(g["seed"]), int(g["page"]), int(g["words"])
    score, geo, _ = synthdata.make_maps(seed, page, words)
    assert sha(score, geo) == str(g["input_sha"])
    quads = ops.decode_quads_from_maps(score, geo.transpose(1, 2, 0), 0.6, 4.0, 2)
    assert len(quads) == int(g["n_candidates"])
    assert sha(quads) == str(g["quads_sha"])
    nms = ops.locality_aware_nms(quads, 0.2)
    np.testing.assert_array_equal(nms, g["lanms_stable"])
    if name != "page_cfg1":  # no interacting x0 tie: equal to the unpatched reference as well
        np.testing.assert_array_equal(nms, g["lanms_ref"])
    orig_hw = tuple(int(v) for v in g["orig_hw"])
    np.testing.assert_array_equal(ops.expand_boxes(nms, 0.9, 0.9), g["expanded"])
    x = ops.east_postprocess(nms, orig_hw, target_size=page)
    np.testing.assert_array_equal(x, g["aligned"])
    x_noalign = ops.east_postprocess(nms, orig_hw, target_size=page, axis_aligned=False)
    np.testing.assert_array_equal(x_noalign, g["anomalies"])
    rects, valid = ops.word_rects(x, orig_hw[0], orig_hw[1], 5)
    np.testing.assert_array_equal(valid, g["valid"])
    np.testing.assert_array_equal(rects[valid], g["rects"][g["valid"]])
    img = synthdata.make_page_image(seed, max(orig_hw))[: orig_hw[0], : orig_hw[1]]
    assert sha(img) == str(g["image_sha"])
    k = len(g["canvas32"])
    batch, canvas = ops.crop_resize_pad(img, rects[valid][:k], 32, 128, want_canvas=True)
    np.testing.assert_array_equal(canvas, g["canvas32"])  # 0 LSB
    ref = ((canvas.astype(np.float32) - np.float32(127.5)) * np.float32(1 / 127.5)).transpose(0, 3, 1, 2)
    np.testing.assert_array_equal(batch, ref)


def test_box_filters_golden(ops, golden_dir):
    g = load(golden_dir, "box_filters")
    kw = dict(target_size=1, expand_w=0.0, expand_h=0.0)
    c = ops.east_postprocess(g["boxes"], (1, 1), axis_aligned=False, remove_anomalies=False, **kw)
    np.testing.assert_array_equal(rows_sorted(c), rows_sorted(g["contained"]))
    np.testing.assert_array_equal(c, cpu.remove_fully_contained_boxes(g["boxes"]))
    a = ops.east_postprocess(g["boxes"], (1, 1), axis_aligned=False, **kw)
    np.testing.assert_array_equal(rows_sorted(a), rows_sorted(g["anomalies"]))
    x = ops.east_postprocess(g["boxes"], (1, 1), **kw)
    np.testing.assert_array_equal(rows_sorted(x), rows_sorted(g["aligned"]))


def test_resize_pad_golden(ops, golden_dir):
    g = load(golden_dir, "resize_pad")
    img = synthdata.make_page_image(int(g["image_seed"]), 512)
    assert sha(img) == str(g["image_sha"])
    rects = g["rects"]
    b32, c32 = ops.crop_resize_pad(img, rects, 32, 128, want_canvas=True)
    c64 = ops.crop_resize_pad(img, rects, 64, 256, want_canvas=True, want_batch=False)
    for i in range(len(rects)):
        assert sha(c32[i]) == str(g["sha32"][i]), f"32x128 case {i} rect {rects[i]}"
        assert sha(c64[i]) == str(g["sha64"][i]), f"64x256 case {i} rect {rects[i]}"
    ref = ((c32.astype(np.float32) - np.float32(127.5)) * np.float32(1 / 127.5)).transpose(0, 3, 1, 2)
    np.testing.assert_array_equal(b32, ref)


# ---- seeded inputs against the oracle -------------------------------------------------------------------------------------
@pytest.mark.parametrize("q", [1, 2, 4])
@pytest.mark.parametrize("seed,page,words", [(11, 384, 60), (12, 640, 200)])
def test_decode_vs_oracle(ops, seed, page, words, q):
    score, geo, _ = synthdata.make_maps(seed, page, words)
    for thr in (0.6, 0.25, 0.95):
        a = ops.decode_quads_from_maps(score, geo, thr, 4.0, q)
        b = cpu.decode_quads_from_maps(score, geo, thr, 4.0, q)
        np.testing.assert_array_equal(a, b)
    a = ops.decode_quads_from_maps(score, geo, 0.6, 3.7, q)  # inexact scale: f64 + f32 promotion path
    np.testing.assert_array_equal(a, cpu.decode_quads_from_maps(score, geo, 0.6, 3.7, q))


def test_decode_edge_cases(ops):
    H = W = 8
    score = np.zeros((H, W), np.float32)
    geo = np.zeros((8, H, W), np.float32)
    assert ops.decode_quads_from_maps(score, geo, 0.6, 4.0, 2).shape == (0, 9)
    score[3, 3] = np.float32(0.6)
    assert len(ops.decode_quads_from_maps(score, geo, 0.6, 4.0, 1)) == 0  # strict >
    score[3, 3] = np.nextafter(np.float32(0.6), np.float32(1))
    q = ops.decode_quads_from_maps(score, geo, 0.6, 4.0, 1)
    assert q.shape == (1, 9) and q[0, 0] == 12.0 and q[0, 1] == 12.0
    score[:] = 0
    score[2, 2] = 0.9
    score[2, 3] = 0.95
    score[3, 3] = 0.1
    q = ops.decode_quads_from_maps(score, geo, 0.6, 4.0, 2)
    assert q.shape == (1, 9) and q[0, 8] == np.float32(0.1)
    s7 = np.zeros((7, 7), np.float32)
    s7[6, 0] = 0.9
    with pytest.raises(IndexError):
        ops.decode_quads_from_maps(s7, np.zeros((8, 7, 7), np.float32), 0.6, 4.0, 2)
    # every pixel a candidate (the random-init case of SURVEY 8c: all 102 400 px -> 25 600 rows)
    full = np.full((64, 96), 0.9, np.float32)
    gg = np.random.default_rng(0).normal(0, 3, (8, 64, 96)).astype(np.float32)
    np.testing.assert_array_equal(ops.decode_quads_from_maps(full, gg, 0.6, 4.0, 2),
                                  cpu.decode_quads_from_maps(full, gg, 0.6, 4.0, 2))
    np.testing.assert_array_equal(ops.decode_quads_from_maps(full, gg, 0.6, 4.0, 1),
                                  cpu.decode_quads_from_maps(full, gg, 0.6, 4.0, 1))


@pytest.mark.parametrize("thr", [0.05, 0.2, 0.5, 0.9])
def test_lanms_vs_oracle(ops, thr):
    for seed, page, words in [(21, 384, 60), (22, 640, 220), (23, 1024, 600)]:
        score, geo, _ = synthdata.make_maps(seed, page, words)
        quads = cpu.decode_quads_from_maps(score, geo, 0.6, 4.0, 2)
        np.testing.assert_array_equal(ops.locality_aware_nms(quads, thr), cpu.locality_aware_nms(quads, thr))


def test_lanms_hard_cases(ops):
    rng = np.random.default_rng(9)
    # (a) one giant cluster: every box overlaps the running merge (longest speculative run)
    n = 400
    base = np.array([10, 10, 60, 10, 60, 30, 10, 30], np.float32)
    boxes = np.tile(base, (n, 1)) + rng.normal(0, 0.3, (n, 8)).astype(np.float32)
    boxes = np.concatenate([boxes, rng.uniform(0.5, 1, (n, 1)).astype(np.float32)], axis=1)
    np.testing.assert_array_equal(ops.locality_aware_nms(boxes, 0.2), cpu.locality_aware_nms(boxes, 0.2))
    # (b) exact duplicates + equal x0 + equal scores (tie rule = original index)
    dup = np.repeat(boxes[:40], 5, axis=0)
    dup[:, 8] = np.float32(0.75)
    dup[::3, :8] += 100
    np.testing.assert_array_equal(ops.locality_aware_nms(dup, 0.3), cpu.locality_aware_nms(dup, 0.3))
    # (c) irregular quads: clockwise, self-intersecting, degenerate, mixed with regular ones
    m = 600
    c = rng.uniform(0, 200, (m, 1, 2))
    quads = c + rng.uniform(-15, 15, (m, 4, 2))
    rect = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], float)
    quads[:300] = c[:300] + rect * rng.uniform(5, 30, (300, 1, 2))
    quads[300:330] = quads[300:330, ::-1]
    quads[330:340] = quads[330:340, :1]  # all four vertices equal
    b2 = np.concatenate([quads.reshape(m, 8), rng.uniform(0.3, 1, (m, 1))], axis=1).astype(np.float32)
    for thr in (0.0, 0.1, 0.3, -1.0):
        np.testing.assert_array_equal(ops.locality_aware_nms(b2, thr), cpu.locality_aware_nms(b2, thr))
    # (d) a single box, two boxes
    np.testing.assert_array_equal(ops.locality_aware_nms(b2[:1], 0.2), cpu.locality_aware_nms(b2[:1], 0.2))
    np.testing.assert_array_equal(ops.locality_aware_nms(b2[:2], 0.2), cpu.locality_aware_nms(b2[:2], 0.2))
    # (e) standard_nms index sets on the irregular mix (exact all-pairs mode)
    keep = ops.standard_nms(quads, b2[:, 8].astype(np.float64), 0.1, return_index=True)
    np.testing.assert_array_equal(keep, cpu.standard_nms(quads, b2[:, 8].astype(np.float64), 0.1, return_index=True))


@pytest.mark.parametrize("thr", [0.0, 0.05, 0.2, 0.5, 0.8])
def test_nms_pairs_around_the_threshold(ops, thr):
    """The resolve step decides most pairs of regular quads by a shrink-and-contain bound before it clips
    (csrc/lanms.cu iou_above_by_containment).  Pairs whose IoU lies on both sides of the threshold, close to it and
    far from it, rotated and of unequal size, arranged so that the two boxes of a pair are never neighbours in the
    x0 order (the merge scan leaves them alone and the NMS has to judge them): same rows as the oracle, bit for bit."""
    rng = np.random.default_rng(int(thr * 100) + 5)
    rows, cols = 24, 40
    rect = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], float)
    boxes = []
    for r in range(rows):
        for c in range(cols):
            w, h = rng.uniform(30, 60), rng.uniform(10, 20)
            org = np.array([c * 90.0 + r * 3.7, r * 40.0])
            a = org + rect * [w, h]
            # second box: shifted / scaled so that the IoU is spread around thr (and sometimes far above it)
            target = np.clip(thr + rng.uniform(-0.15, 0.15), 0.0, 0.98) if rng.random() < 0.7 else rng.uniform(0.6, 0.98)
            shift = w * (1 - target) / (1 + target)  # IoU of two equal rectangles shifted along x
            sc = rng.uniform(0.9, 1.1)
            b = org + [shift, rng.uniform(-0.5, 0.5)] + rect * [w * sc, h * sc]
            for qd in (a, b):
                th = rng.uniform(-0.06, 0.06)
                rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
                ctr = qd.mean(0)
                boxes.append(np.concatenate([((qd - ctr) @ rot.T + ctr).ravel(), [rng.uniform(0.3, 1.0)]]))
    b = np.asarray(boxes, np.float32)
    b = b[rng.permutation(len(b))]
    got, want = ops.locality_aware_nms(b, thr), cpu.locality_aware_nms(b, thr)
    assert 0 < len(want) < len(b)
    np.testing.assert_array_equal(got, want)


def test_expand_and_filters_vs_oracle(ops):
    for seed, page, words, orig in [(31, 640, 200, (900, 700)), (32, 1024, 600, (1024, 1024))]:
        score, geo, _ = synthdata.make_maps(seed, page, words)
        nms = cpu.locality_aware_nms(cpu.decode_quads_from_maps(score, geo, 0.6, 4.0, 2), 0.2)
        for ew, eh in [(0.9, 0.9), (0.3, 0.7), (0.0, 0.5), (0.0, 0.0)]:
            np.testing.assert_array_equal(ops.expand_boxes(nms, ew, eh), cpu.expand_boxes(nms, ew, eh))
            for aa in (True, False):
                got = ops.east_postprocess(nms, orig, target_size=page, expand_w=ew, expand_h=eh, axis_aligned=aa)
                want = cpu.east_postprocess(nms, orig, target_size=page, expand_w=ew, expand_h=eh, axis_aligned=aa)
                np.testing.assert_array_equal(got, want)
    # nested boxes so that contained-box removal and the anomaly filter both fire
    rng = np.random.default_rng(4)
    n = 300
    cxy = rng.uniform(50, 950, (n, 2))
    wh = rng.uniform(10, 60, (n, 2))
    wh[:5] *= 12  # area anomalies
    rect = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], float) / 2
    q = (cxy[:, None, :] + rect[None] * wh[:, None, :]).reshape(n, 8)
    inner = q[:60].reshape(60, 4, 2)
    inner = inner.mean(axis=1, keepdims=True) + (inner - inner.mean(axis=1, keepdims=True)) * 0.5
    boxes = np.concatenate([np.concatenate([q, inner.reshape(60, 8)]), rng.uniform(0.5, 1, (n + 60, 1))], axis=1)
    boxes = boxes.astype(np.float32)
    got = ops.east_postprocess(boxes, (1000, 1000), target_size=1000, expand_w=0.0, expand_h=0.0)
    want = cpu.east_postprocess(boxes, (1000, 1000), target_size=1000, expand_w=0.0, expand_h=0.0)
    np.testing.assert_array_equal(got, want)
    assert len(want) < len(boxes) - 30


def test_word_rects_and_crops_vs_oracle(ops):
    rng = np.random.default_rng(8)
    H, W = 300, 420
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    n = 400
    xy = rng.uniform(-30, 450, (n, 4, 2)).astype(np.float32)
    xy[:, :, 1] = rng.uniform(-30, 330, (n, 4)).astype(np.float32)
    small = rng.uniform(0, 400, (n // 2, 1, 2)) + rng.uniform(0, 40, (n // 2, 4, 2))
    xy[: n // 2] = small.astype(np.float32)
    r_got, v_got = ops.word_rects(xy, H, W, 5)
    r_want, v_want = cpu.word_rects(xy, H, W, 5)
    np.testing.assert_array_equal(v_got, v_want)
    np.testing.assert_array_equal(r_got[v_got], r_want[v_want])
    assert 20 < v_want.sum() < n
    rects = r_want[v_want]
    for ih, iw in [(32, 128), (64, 256), (48, 100)]:
        batch, canvas = ops.crop_resize_pad(img, rects, ih, iw, want_canvas=True)
        for i, r in enumerate(rects):
            c, chw = cpu.crop_resize_pad(img, r, ih, iw)
            np.testing.assert_array_equal(canvas[i], c, err_msg=f"rect {r} -> {ih}x{iw}")
            np.testing.assert_array_equal(batch[i], chw)


def test_errors_are_loud(ops):
    with pytest.raises(ValueError):
        ops.decode_quads_from_maps(np.zeros((4, 4), np.float32), np.zeros((7, 4, 4), np.float32), 0.5, 4.0)
    with pytest.raises(ValueError):
        ops.locality_aware_nms(np.zeros((3, 8), np.float32), 0.2)


def test_lanms_dense_candidates(ops):
    """Far more overlapping neighbours per box than the default pair capacity (16 per candidate, 32 per cluster):
    the exact two-pass rebuild and the capacity retry of the host entry points must give the oracle's answer."""
    rng = np.random.default_rng(17)
    n = 3000
    rect = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], float)
    c = rng.uniform(0, 260, (n, 1, 2))
    quads = c + rect * rng.uniform(15, 45, (n, 1, 2))
    boxes = np.concatenate([quads.reshape(n, 8), rng.uniform(0.3, 1, (n, 1))], axis=1).astype(np.float32)
    for thr in (0.2, 0.6):
        np.testing.assert_array_equal(ops.locality_aware_nms(boxes, thr), cpu.locality_aware_nms(boxes, thr))
    # every quantisation cell a candidate (score threshold below the background): background geometry is
    # zero there, so most quads are degenerate points -- the irregular all-pairs path
    score, geo, _ = synthdata.make_maps(7, 512, 80)
    q = cpu.decode_quads_from_maps(score, geo, -1.0, 4.0, 2)
    assert len(q) == 64 * 64
    np.testing.assert_array_equal(ops.decode_quads_from_maps(score, geo, -1.0, 4.0, 2), q)
    np.testing.assert_array_equal(ops.locality_aware_nms(q, 0.2), cpu.locality_aware_nms(q, 0.2))


# ---- rectified crops of rotated quads (an extension; cv2 is the specification) ------------------------
def test_quad_warp_golden(ops, golden_dir):
    """ms_warp_quad_host against cv2.warpPerspective's own output, both border modes, bit for bit."""
    g = np.load(os.path.join(golden_dir, "quad_warp.npz"))
    rng = np.random.default_rng(int(g["seed"]))
    H, W = (int(v) for v in g["page_hw"])
    page = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    off = 0
    bv = int(g["border_value"])
    for q, (w, h) in zip(g["quads"], g["sizes"]):
        got = ops.warp_quad(page, q, "constant", bv)
        if w == 0:
            assert got is None
            continue
        n = int(w) * int(h) * 3
        assert got.shape == (h, w, 3)
        np.testing.assert_array_equal(got.reshape(-1), g["const"][off:off + n])
        np.testing.assert_array_equal(ops.warp_quad(page, q, "replicate").reshape(-1), g["repl"][off:off + n])
        off += n
    assert off == len(g["const"])


def _random_quads(rng, n, H, W, wmax, hmax):
    out = []
    for _ in range(n):
        cx, cy = rng.uniform(0, W), rng.uniform(0, H)
        ww, hh = rng.uniform(3, wmax), rng.uniform(3, hmax)
        ang = rng.uniform(-0.7, 0.7)
        c, s = np.cos(ang), np.sin(ang)
        q = np.array([[-ww / 2, -hh / 2], [ww / 2, -hh / 2], [ww / 2, hh / 2], [-ww / 2, hh / 2]]) @ np.array(
            [[c, s], [-s, c]]) + [cx, cy]
        out.append((q + rng.uniform(-2, 2, q.shape)).reshape(-1))
    return np.array(out, np.float32)


@pytest.mark.parametrize("out_hw", [(32, 128), (64, 256)])
def test_quad_crop_vs_oracle(ops, out_hw):
    """Warp + ResizeAndPadA + normalise against the oracle chain: small quads (INTER_LINEAR upscale), typical word
    quads (INTER_AREA), patches larger than the shared-memory stage (evaluated on the fly), quads hanging over the
    page border, quads below min_text_size and degenerate ones (no patch -> all-padding row, valid False)."""
    ih, iw = out_hw
    rng = np.random.default_rng(77)
    H, W = 400, 700
    page = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    quads = np.concatenate([
        _random_quads(rng, 40, H, W, 30, 12),      # small: upscale / below min_text_size
        _random_quads(rng, 60, H, W, 200, 60),     # word-sized
        _random_quads(rng, 6, H, W, 600, 90)[:, :],  # larger than the 40 KB stage
        np.array([[5, 5, 5.2, 5, 5.2, 5.2, 5, 5.2], [10, 20, 74, 20, 74, 36, 10, 36],
                  [np.nan, 0, 50, 0, 50, 20, 0, 20], [0, 0, 3e38, 0, 3e38, 20, 0, 20]], np.float32),
    ])
    for border, bv in (("constant", 0), ("replicate", 0), ("constant", 200)):
        batch, canvas, valid = ops.quad_crop_resize_pad(page, quads, ih, iw, 5, border, bv, want_canvas=True)
        n_valid = 0
        for i, q in enumerate(quads):
            want_canvas, want_chw = cpu.quad_crop_resize_pad(page, q, ih, iw, 5, border, bv)
            if want_canvas is None:
                assert not valid[i]
                assert (canvas[i] == 255).all() and (batch[i] == 1.0).all()
                continue
            assert valid[i], (i, q)
            n_valid += 1
            np.testing.assert_array_equal(canvas[i], want_canvas, err_msg=f"quad {i} {border}")
            np.testing.assert_array_equal(batch[i], want_chw)
        assert n_valid >= 80 and not valid[-1] and not valid[-2] and not valid[-4] and valid[-3]
    # patch alone, larger than one launch wave of the single-quad kernel
    big = np.array([20, 30, 660, 60, 650, 360, 15, 330], np.float32)
    np.testing.assert_array_equal(ops.warp_quad(page, big, "replicate"), cpu.warp_quad(page, big, "replicate"))


def _quad_counts(ctx):
    import ctypes as C

    c = (C.c_int32 * 2)()
    from manuscript_b200 import _cabi

    _cabi.check(ctx.lib.ms_quad_crop_last_counts(ctx.handle, c))
    return int(c[0]), int(c[1])


def test_quad_crop_staged_kernel(ops):
    """The staged rotated-crop kernel (TMA window + per-warp bands + dp4a blend) takes a quad when its window lies
    inside a page whose rows are 16-byte aligned.  Checked (a) against the oracle chain (cv2-pinned) quad by quad, float32
    batch + uint8 canvas, (b) against the generic kernel (MS_B200_QUAD_NO_STAGE=1) on thousands of quads -- turned word
    boxes, strong perspective, tall / narrow / tiny quads, quads over the page border (generic kernel), non-convex
    quads (handed back) -- bit for bit, both output variants."""
    import manuscript_b200 as mb

    rng = np.random.default_rng(123)
    H, W = 416, 704  # 704 * 3 bytes per row: a multiple of 16
    page = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    quads = np.concatenate([
        _random_quads(rng, 120, H, W, 200, 60),   # word-sized, turned by up to +-0.7 rad
        _random_quads(rng, 30, H, W, 30, 12),     # small
        _random_quads(rng, 10, H, W, 380, 90),    # wide
        np.array([[100, 100, 220, 90, 200, 160, 90, 150],     # perspective
                  [300, 200, 420, 200, 420, 240, 300, 240],   # axis aligned
                  [300, 100, 400, 100, 320, 130, 300, 160],   # concave: taps leave the hull -> handed back
                  [50, 300, 150, 300.5, 150, 340, 50, 340.5],
                  [np.nan, 0, 50, 0, 50, 20, 0, 20], [0, 0, 3e38, 0, 3e38, 20, 0, 20],     # no patch
                  [5, 5, 5.2, 5, 5.2, 5.2, 5, 5.2], [10, 20, 13, 20, 13, 60, 10, 60],     # degenerate / below min_text_size
                  [100, 100, 200, 100, 100, 100.001, 200, 100.001],                       # nearly collinear
                  [-20, -10, 90, -12, 92, 30, -18, 28], [650, 380, 720, 384, 718, 420, 648, 416]], np.float32),  # over the border
    ])
    ctx = mb.ops.default_context()
    batch, canvas, valid = ops.quad_crop_resize_pad(page, quads, 32, 128, 5, "constant", 7, want_canvas=True)
    n_staged, n_generic = _quad_counts(ctx)
    assert n_staged >= 25 and n_generic >= 20, (n_staged, n_generic)
    for i, q in enumerate(quads):
        want_canvas, want_chw = cpu.quad_crop_resize_pad(page, q, 32, 128, 5, "constant", 7)
        if want_canvas is None:
            assert not valid[i] and (canvas[i] == 255).all()
            continue
        np.testing.assert_array_equal(canvas[i], want_canvas, err_msg=f"quad {i}")
        np.testing.assert_array_equal(batch[i], want_chw, err_msg=f"quad {i}")
    only_f32, _ = ops.quad_crop_resize_pad(page, quads, 32, 128, 5, "constant", 7)
    np.testing.assert_array_equal(only_f32, batch)
    # a larger population against the generic kernel
    os.environ["MS_B200_QUAD_NO_STAGE"] = "1"
    try:
        gctx = mb._cabi.Context(0)
    finally:
        del os.environ["MS_B200_QUAD_NO_STAGE"]
    H2, W2 = 1024, 1024
    page2 = synthdata.make_page_image(5, H2)
    many = np.concatenate([_random_quads(rng, 3000, H2, W2, 160, 70), _random_quads(rng, 500, H2, W2, 60, 25)])
    many[::7] += rng.uniform(-6, 6, many[::7].shape).astype(np.float32)  # stronger perspective
    for border in ("constant", "replicate"):
        a_f, a_u, a_v = ops.quad_crop_resize_pad(page2, many, 32, 128, 5, border, 0, want_canvas=True)
        ns, ng = _quad_counts(ctx)
        b_f, b_u, b_v = ops.quad_crop_resize_pad(page2, many, 32, 128, 5, border, 0, want_canvas=True, ctx=gctx)
        assert _quad_counts(gctx)[0] == 0
        assert ns > 1000, (ns, ng)
        np.testing.assert_array_equal(a_v, b_v)
        np.testing.assert_array_equal(a_u, b_u)
        np.testing.assert_array_equal(a_f, b_f)
    gctx.close()


# ---- the NMS shortcut predicates themselves (test-only C-ABI entry ms_test_iou_proved_host) ---------------------
def _rot(q, th, ctr=None):
    ctr = q.mean(1, keepdims=True) if ctr is None else ctr
    rot = np.stack([np.stack([np.cos(th), -np.sin(th)], -1), np.stack([np.sin(th), np.cos(th)], -1)], -2)
    return np.einsum("nij,nkj->nki", rot, q - ctr) + ctr


def _predicate_pairs(rng, n, thr):
    """Four families of quad pairs, n each, (n,4,2) float64, positively oriented unless stated."""
    rect = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], float)
    out = []
    # A: unrelated sizes / shifts / rotations anywhere on pages up to 16384 px
    w, h = rng.uniform(4, 400, (n, 1, 1)), rng.uniform(3, 120, (n, 1, 1))
    wh = np.concatenate([w, h], axis=2)
    org = rng.uniform(0, 16384, (n, 1, 2))
    a = rect[None] * wh + org + rng.normal(0, 0.02, (n, 4, 2)) * wh
    frac = rng.uniform(-1.0, 1.0, (n, 1, 2)) * rng.choice([0.02, 0.2, 0.6, 1.0], (n, 1, 1))
    b = rect[None] * wh * rng.uniform(0.5, 1.6, (n, 1, 1)) + org + frac * wh + rng.normal(0, 0.02, (n, 4, 2)) * wh
    out.append((_rot(a, rng.uniform(-0.4, 0.4, n)), _rot(b, rng.uniform(-0.4, 0.4, n))))
    # B: slivers -- parallelograms whose corner angle has |sin| between 3e-4 and 5e-3 (the regularity bound is 1e-3),
    #    and edges between 3e-4 and 5e-3 px (bound 1e-3); the partner is the same quad moved a little
    s = 10.0 ** rng.uniform(-3.5, -2.3, n)
    L = rng.uniform(10, 300, n)
    H = np.where(rng.random(n) < 0.5, rng.uniform(5, 60, n), 10.0 ** rng.uniform(-3.5, -2.3, n))
    ang = np.arcsin(np.clip(s, 0, 1)) * np.where(rng.random(n) < 0.7, 1.0, 300.0)  # some clearly regular
    u = np.stack([L, np.zeros(n)], -1)
    v = np.stack([H * np.cos(ang), H * np.sin(ang)], -1)
    org = rng.uniform(0, 16384, (n, 2))
    a = np.stack([org, org + u, org + u + v, org + v], axis=1)
    b = a + rng.normal(0, 1.0, (n, 1, 2)) * 10.0 ** rng.uniform(-4, 0.5, (n, 1, 1))
    th = rng.uniform(-3.1, 3.1, n)
    out.append((_rot(a, th), _rot(b, th, a.mean(1, keepdims=True))))
    # C: candidates of one word -- the same quad plus decode noise, at large coordinates
    w, h = rng.uniform(20, 300, (n, 1, 1)), rng.uniform(8, 80, (n, 1, 1))
    wh = np.concatenate([w, h], axis=2)
    org = rng.uniform(0, 16384, (n, 1, 2))
    a = _rot(rect[None] * wh + org, rng.uniform(-0.1, 0.1, n))
    b = a + rng.normal(0, 1.0, (n, 4, 2)) * 10.0 ** rng.uniform(-2, 0.7, (n, 1, 1))
    out.append((a, b))
    # D: IoU steered to the threshold: equal rectangles shifted along x by w(1-t)/(1+t), t = thr +- tiny
    w, h = rng.uniform(20, 300, (n, 1, 1)), rng.uniform(8, 80, (n, 1, 1))
    wh = np.concatenate([w, h], axis=2)
    t = np.clip(thr + rng.normal(0, 1.0, n) * 10.0 ** rng.uniform(-6, -1, n), 1e-9, 0.999)
    shift = w[:, 0, 0] * (1 - t) / (1 + t)
    org = rng.uniform(0, 16384, (n, 1, 2))
    a = rect[None] * wh + org
    b = a + np.stack([shift, np.zeros(n)], -1)[:, None, :]
    th = rng.uniform(-0.2, 0.2, n)
    out.append((_rot(a, th), _rot(b, th, a.mean(1, keepdims=True))))  # one rigid motion for the pair
    return out


@pytest.mark.parametrize("thr", [0.0, 0.2, 0.8])
def test_containment_predicate_on_the_device(ops, thr):
    """csrc/lanms.cu iou_above_by_containment / quad_regular_bbox, evaluated BY THE DEVICE FUNCTIONS on 1.4 million
    pairs per threshold (coordinates up to 16384, slivers on both sides of the 1e-3 regularity bounds, same-word
    candidates, IoU steered onto the threshold): "proven above" never contradicts the oracle's float64
    Sutherland-Hodgman IoU (lanms.py:80-96), and the shortcut still decides most same-word pairs."""
    import ctypes as C

    from manuscript_b200._cabi import check, default_context

    rng = np.random.default_rng(int(thr * 100) + 77)
    n = 350_000
    ctx = default_context()
    total = proved_total = 0
    for fam, (a, b) in zip("ABCD", _predicate_pairs(rng, n, thr)):
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        out = np.zeros(len(a), np.uint8)
        check(ctx.lib.ms_test_iou_proved_host(ctx.handle, a.ctypes.data, b.ctypes.data, len(a), float(thr),
                                              out.ctypes.data))
        iou = cpu.polygon_iou_batch(a, b)
        regular, proved = (out & 1) != 0, (out & 2) != 0
        assert not np.any(proved & ~regular), fam
        bad = proved & ~(iou > thr)
        assert not bad.any(), (fam, thr, int(bad.sum()), a[bad][:1], b[bad][:1], iou[bad][:1])
        if fam == "B":  # both sides of the regularity bound are present
            assert 0.05 < regular.mean() < 0.95, regular.mean()
        if fam == "C" and thr < 0.5:  # (for thr >= ~0.74 the bound would need lam^2 >= 0.9 and is never tried)
            clear = iou > min(0.97, thr + 0.4 * (1 - thr) + 0.1)
            assert clear.sum() > 1000 and proved[clear].mean() > 0.6, (clear.sum(), proved[clear].mean())
        if fam == "D":  # on the threshold the bound must stay silent (it promises a 6 % margin)
            assert not np.any(proved & (iou <= thr * 1.02 + 1e-12))
        total += len(a)
        proved_total += int(proved.sum())
    assert total >= 1_000_000 and (proved_total > 100_000 or thr >= 0.5)


def test_context_is_guarded_against_a_second_thread(ops):
    """An ms_ctx owns one scratch arena and is single-threaded by contract; a second thread that enters while a call is
    in flight gets MS_ERR_INVALID ("in use by another thread") instead of corrupting the arena, and the first call's
    result is untouched.  Each thread's own default context is unaffected."""
    import threading

    from manuscript_b200._cabi import Context, last_error

    ctx = Context(0)
    score, geo, _ = synthdata.make_maps(11, 2048, 2000)
    quads = ops.decode_quads_from_maps(score, geo, 0.6, 4.0, 2, ctx=ctx)
    want = cpu.locality_aware_nms(quads, 0.2)
    seen = {"busy": 0, "ok": 0, "other": []}
    stop = threading.Event()

    def intruder():
        a = np.zeros((4, 4, 2), np.float64)
        out = np.zeros(4, np.float64)
        while not stop.is_set():
            rc = ctx.lib.ms_polygon_iou_host(ctx.handle, a.ctypes.data, a.ctypes.data, 4, out.ctypes.data)
            if rc == 0:
                seen["ok"] += 1
            elif rc == -1 and "another thread" in last_error():
                seen["busy"] += 1
            else:
                seen["other"].append((rc, last_error()))

    t = threading.Thread(target=intruder)
    t.start()
    done = turned_away = 0
    try:
        while done < 20:
            try:
                got = ops.locality_aware_nms(quads, 0.2, ctx=ctx)
            except ops.CABIError as e:  # the intruder was inside its (short) call: this thread is the second one
                assert e.code == -1 and "another thread" in str(e)
                turned_away += 1
                continue
            np.testing.assert_array_equal(got, want)  # a call that was let in is never disturbed
            done += 1
    finally:
        stop.set()
        t.join()
    assert not seen["other"], seen["other"][:3]
    assert seen["busy"] + turned_away > 0  # the two threads did collide, and the second one was turned away each time
