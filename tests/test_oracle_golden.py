"""Pin the C oracle against golden vectors produced by the real reference
(tests/golden/make_golden.py) -- bit-exact everywhere (integer/byte/index work and the f32/f64
arithmetic the reference performs)."""
import hashlib
import os

import numpy as np
import pytest

import synthdata
from oracle import cpu, refload

PAGES = ["page_s0", "page_s1", "page_s2", "page_cfg1"]


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("name", PAGES)
def test_page_chain(golden_dir, name):
    g = load(golden_dir, name)
    seed, page, words = int(g["seed"]), int(g["page"]), int(g["words"])
    score, geo, _ = synthdata.make_maps(seed, page, words)
    assert sha(score, geo) == str(g["input_sha"]), "synthetic generator drifted; regenerate golden"
    quads = cpu.decode_quads_from_maps(score, geo, 0.6, 4.0, 2)
    assert len(quads) == int(g["n_candidates"])
    assert sha(quads) == str(g["quads_sha"])
    np.testing.assert_array_equal(quads[:64], g["quads_head"])
    nms = cpu.locality_aware_nms(quads, 0.2)
    np.testing.assert_array_equal(nms, g["lanms_stable"])
    # vs the unpatched reference: same multiset of rows whenever its unstable sort met no interacting tie
    same = nms.shape == g["lanms_ref"].shape and np.array_equal(nms, g["lanms_ref"])
    if name != "page_cfg1":
        assert same
    orig_hw = tuple(int(v) for v in g["orig_hw"])
    e = cpu.expand_boxes(nms, 0.9, 0.9)
    np.testing.assert_array_equal(e, g["expanded"])
    s = cpu.scale_boxes_to_original(e, orig_hw, page)
    np.testing.assert_array_equal(s, g["scaled"])
    c = cpu.remove_fully_contained_boxes(s)
    np.testing.assert_array_equal(c, g["contained"])
    a = cpu.remove_area_anomalies(c)
    np.testing.assert_array_equal(a, g["anomalies"])
    x = cpu.convert_to_axis_aligned(a)
    np.testing.assert_array_equal(x, g["aligned"])
    rects, valid = cpu.word_rects(x, orig_hw[0], orig_hw[1], 5)
    np.testing.assert_array_equal(valid, g["valid"])
    np.testing.assert_array_equal(rects[valid], g["rects"][g["valid"]])
    img = synthdata.make_page_image(seed, max(orig_hw))[: orig_hw[0], : orig_hw[1]]
    assert sha(img) == str(g["image_sha"])
    k = 0
    for r, ok in zip(rects, valid):
        if not ok or k >= len(g["canvas32"]):
            continue
        canvas, chw = cpu.crop_resize_pad(img, r, 32, 128)
        np.testing.assert_array_equal(canvas, g["canvas32"][k])
        k += 1
    assert k == len(g["canvas32"])


def test_cfg1_tie_deviation_is_small(golden_dir):
    """page_cfg1 has x0 ties that the reference's unstable argsort orders differently from the
    stable rule; the deviation is a handful of rows and is recorded, not hidden."""
    g = load(golden_dir, "page_cfg1")
    a = set(map(bytes, g["lanms_ref"]))
    b = set(map(bytes, g["lanms_stable"]))
    assert len(g["lanms_ref"]) == len(g["lanms_stable"]) == 500
    assert len(a ^ b) <= 12


def rows_sorted(a):
    a = np.ascontiguousarray(a)
    return a[np.lexsort(a.T[::-1])]


def test_box_filters(golden_dir):
    """The fixture holds exact duplicate boxes (equal areas): which twin survives depends on the
    reference's unstable argsort (infer.py:199), so rows are compared as a multiset."""
    g = load(golden_dir, "box_filters")
    c = cpu.remove_fully_contained_boxes(g["boxes"])
    np.testing.assert_array_equal(rows_sorted(c), rows_sorted(g["contained"]))
    a = cpu.remove_area_anomalies(c)
    np.testing.assert_array_equal(rows_sorted(a), rows_sorted(g["anomalies"]))
    assert len(a) < len(c) < len(g["boxes"])
    np.testing.assert_array_equal(rows_sorted(cpu.convert_to_axis_aligned(a)), rows_sorted(g["aligned"]))


def test_point_polygon(golden_dir):
    g = load(golden_dir, "point_polygon")
    got = [cpu.point_polygon_test(c, p) for c, p in zip(g["contours"], g["points"])]
    np.testing.assert_array_equal(np.array(got, np.int8), g["result"])
    assert set(np.unique(g["result"])) == {-1, 0, 1}


def test_resize_pad(golden_dir):
    g = load(golden_dir, "resize_pad")
    img = synthdata.make_page_image(int(g["image_seed"]), 512)
    assert sha(img) == str(g["image_sha"])
    modes = set()
    for i, r in enumerate(g["rects"]):
        c32, chw = cpu.crop_resize_pad(img, r, 32, 128)
        c64, _ = cpu.crop_resize_pad(img, r, 64, 256)
        assert sha(c32) == str(g["sha32"][i]), f"32x128 case {i} rect {r}"
        assert sha(c64) == str(g["sha64"][i]), f"64x256 case {i} rect {r}"
        if i < len(g["canvas32"]):
            np.testing.assert_array_equal(c32, g["canvas32"][i])
        modes.add(cpu.resize_plan(r[3] - r[1], r[2] - r[0], 32, 128)["interp"])
        # normalisation restated from albumentations' formula: (x-127.5)*(1/127.5), HWC->CHW
        ref = ((c32.astype(np.float32) - np.float32(127.5)) * np.float32(1 / 127.5)).transpose(2, 0, 1)
        np.testing.assert_array_equal(chw, ref)
    assert modes == {"copy", "linear", "area"}


def test_np_sum_model():
    rng = np.random.default_rng(3)
    for n in list(range(1, 200)) + [1000, 4097, 20000]:
        v = (rng.standard_normal(n) * 1000).astype(np.float32)
        assert np.sum(v) == cpu.lib().orc_np_sum_f32(v, n)


def test_decode_edge_cases():
    H = W = 8
    score = np.zeros((H, W), np.float32)
    geo = np.zeros((8, H, W), np.float32)
    assert cpu.decode_quads_from_maps(score, geo, 0.6, 4.0, 2).shape == (0, 9)
    # strict > in float32: f32(0.6) is not > 0.6
    score[3, 3] = np.float32(0.6)
    assert len(cpu.decode_quads_from_maps(score, geo, 0.6, 4.0, 1)) == 0
    score[3, 3] = np.nextafter(np.float32(0.6), np.float32(1))
    q = cpu.decode_quads_from_maps(score, geo, 0.6, 4.0, 1)
    assert q.shape == (1, 9) and q[0, 0] == 12.0 and q[0, 1] == 12.0
    # quantisation reads the quantised pixel (here (3,3) for cell (1,1)) and dedups the cell
    score[:] = 0
    score[2, 2] = 0.9
    score[2, 3] = 0.95
    score[3, 3] = 0.1
    q = cpu.decode_quads_from_maps(score, geo, 0.6, 4.0, 2)
    assert q.shape == (1, 9) and q[0, 8] == np.float32(0.1)  # score of the quantised pixel, below thr
    # odd map with q=2: a hit in the last row quantises outside the map -> IndexError like the reference
    s7 = np.zeros((7, 7), np.float32)
    s7[6, 0] = 0.9
    with pytest.raises(IndexError):
        cpu.decode_quads_from_maps(s7, np.zeros((8, 7, 7), np.float32), 0.6, 4.0, 2)


@pytest.mark.skipif(not refload.available(), reason="live reference only in the build container")
def test_live_reference_random_pages():
    """Extra seeds against the live reference (stable-patched), decode -> lanms -> expand."""
    ru, rs = refload.utils(), refload.lanms_stable()
    for seed in (21, 22):
        score, geo, _ = synthdata.make_maps(seed, 384, 60)
        for q in (1, 2, 4):
            a = cpu.decode_quads_from_maps(score, geo, 0.6, 4.0, q)
            b = ru.decode_quads_from_maps(score, geo.transpose(1, 2, 0), 0.6, 4.0, q)
            np.testing.assert_array_equal(a, b)
        for thr in (0.05, 0.2, 0.5):
            np.testing.assert_array_equal(cpu.locality_aware_nms(a, thr), rs.locality_aware_nms(a, thr))
        np.testing.assert_array_equal(cpu.expand_boxes(a[:50], 0.3, 0.7), ru.expand_boxes(a[:50], 0.3, 0.7))


def _quad_warp_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "quad_warp.npz"))
    rng = np.random.default_rng(int(g["seed"]))
    H, W = (int(v) for v in g["page_hw"])
    return g, rng.integers(0, 256, (H, W, 3), dtype=np.uint8)


def test_quad_warp_golden(golden_dir):
    """SURVEY 8f-4 (an extension, no reference function): the oracle's getPerspectiveTransform / warpPerspective
    restatement against cv2's own output (tests/golden/make_golden_quad_warp.py), bit for bit."""
    g, page = _quad_warp_golden(golden_dir)
    off = 0
    n_patches = 0
    for q, (w, h), m in zip(g["quads"], g["sizes"], g["mats"]):
        assert cpu.quad_patch_size(q) == (w, h)
        if w == 0:
            assert cpu.warp_quad(page, q) is None
            assert cpu.quad_crop_resize_pad(page, q, 32, 128) == (None, None)
            continue
        rect = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], np.float32)
        mine = cpu.perspective_transform(rect, q.reshape(4, 2))
        np.testing.assert_array_equal(mine.view(np.uint64), m.view(np.uint64))
        n = int(w) * int(h) * 3
        bv = int(g["border_value"])
        np.testing.assert_array_equal(cpu.warp_quad(page, q, "constant", bv).reshape(-1), g["const"][off:off + n])
        np.testing.assert_array_equal(cpu.warp_quad(page, q, "replicate").reshape(-1), g["repl"][off:off + n])
        off += n
        n_patches += 1
    assert off == len(g["const"]) and n_patches >= 40


def test_large_pages_cfg2_reference_and_tie_deviation(golden_dir):
    """BASELINE configs[2] (2048x2048, ~2000 words, seed 0 -- page 0 of bench.py's batch): tests/golden/
    make_golden_large.py ran the REAL reference on it.  The oracle equals the reference with the stable tie rule
    bit for bit; against the unpatched reference (numpy's unstable argsort, 137 x0 ties among 15 190 candidates)
    the kept sets differ in 4 of 2000 + 2000 rows -- the documented deviation, quantified at the benchmarked shape."""
    g = load(golden_dir, "large_pages")
    page, words, seed = int(g["cfg2_page"]), int(g["cfg2_words"]), int(g["cfg2_seed"])
    score, geo, _ = synthdata.make_maps(seed, page, words)
    assert sha(score, geo) == str(g["cfg2_input_sha"]), "synthetic generator drifted; regenerate golden"
    quads = cpu.decode_quads_from_maps(score, geo, 0.6, 4.0, 2)
    assert len(quads) == int(g["cfg2_n_candidates"]) and sha(quads) == str(g["cfg2_quads_sha"])
    nms = cpu.locality_aware_nms(quads, 0.2)
    np.testing.assert_array_equal(nms, g["cfg2_lanms_stable"])
    assert sha(nms) == str(g["cfg2_lanms_stable_sha"])
    a, b = set(map(bytes, g["cfg2_lanms_ref"])), set(map(bytes, g["cfg2_lanms_stable"]))
    assert len(g["cfg2_lanms_ref"]) == len(g["cfg2_lanms_stable"]) == 2000
    assert len(a ^ b) == int(g["cfg2_rows_differing"]) == 4 and int(g["cfg2_x0_ties"]) == 137
    final = cpu.east_postprocess(nms, (page, page), target_size=page)
    assert sha(final) == str(g["cfg2_final_sha"]) and len(final) == int(g["cfg2_n_final"])
    rects, valid = cpu.word_rects(final, page, page, 5)
    assert sha(rects[valid].astype(np.int32)) == str(g["cfg2_rects_sha"])
    img = synthdata.make_page_image(seed, page)
    assert sha(img) == str(g["cfg2_image_sha"])
    for k, r in enumerate(rects[valid][::40]):
        canvas, chw = cpu.crop_resize_pad(img, r, 32, 128)
        np.testing.assert_array_equal(canvas, g["cfg2_canvas_every40"][k])


def test_large_pages_cfg3_digests(golden_dir):
    """BASELINE configs[3] (4096x4096, ~10 000 words, 63 304 candidates): digests of the oracle's candidates, kept rows
    and final boxes, and the reading order the REAL reference gives them (utils.py:610-644), committed once so that
    the GPU test can assert them without the O(n^2) run.  This test re-derives them (about a minute)."""
    g = load(golden_dir, "large_pages")
    page, words, seed = int(g["cfg3_page"]), int(g["cfg3_words"]), int(g["cfg3_seed"])
    score, geo, _ = synthdata.make_maps(seed, page, words)
    assert sha(score, geo) == str(g["cfg3_input_sha"])
    quads = cpu.decode_quads_from_maps(score, geo, 0.6, 4.0, 2)
    assert len(quads) == int(g["cfg3_n_candidates"]) and sha(quads) == str(g["cfg3_quads_sha"])
    nms = cpu.locality_aware_nms(quads, 0.2)
    assert len(nms) == int(g["cfg3_n_kept"]) and sha(nms) == str(g["cfg3_lanms_sha"])
    final = cpu.east_postprocess(nms, (page, page), target_size=page)
    assert sha(final) == str(g["cfg3_final_sha"])
    # the exact host restatement of the reading order (manuscript_b200.reading_order, no device needed) agrees with
    # the reference's order on all 10 000 boxes
    from manuscript_b200 import reading_order as ro

    keys = [tuple(int(v) for v in ro.int_bbox(row[:8].reshape(4, 2))) for row in final]
    first = {}
    for i, k in enumerate(keys):
        first.setdefault(k, i)
    order = np.array([first[tuple(int(v) for v in bx)] for bx in ro.sort_boxes_reading_order_with_resolutions(keys)],
                     np.int32)
    np.testing.assert_array_equal(order, g["cfg3_order"])
    assert sha(order) == str(g["cfg3_order_sha"])
