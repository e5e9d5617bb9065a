"""Known-answer tests of the reference (tests/detectors/east/test_lanms.py:18-188 in the reference
tree) restated against the C oracle.  The same cases run against the CUDA path in
test_gpu_lanms.py."""
import numpy as np
import pytest

from oracle import cpu

SQ4 = np.array([[0, 0], [4, 0], [4, 4], [0, 4]], dtype=np.float64)
SQ4B = np.array([[2, 2], [6, 2], [6, 6], [2, 6]], dtype=np.float64)
UNIT = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=np.float64)


def test_area_square_triangle_degenerate():
    assert cpu.polygon_area(UNIT) == pytest.approx(1.0, rel=1e-5)
    assert cpu.polygon_area(np.array([[0, 0], [2, 0], [0, 2]], float)) == pytest.approx(2.0, rel=1e-5)
    assert cpu.polygon_area(np.array([[0, 0], [1, 0]], float)) == pytest.approx(0.0)


def test_line_hit_and_parallel():
    hit = cpu.compute_intersection([0, 0], [2, 2], [0, 2], [2, 0])
    np.testing.assert_allclose(hit, [1, 1], rtol=1e-5)
    par = cpu.compute_intersection([0, 0], [1, 1], [2, 2], [3, 3])
    np.testing.assert_allclose(par, [0, 0], rtol=1e-5)


def test_clip_halfplane_cases():
    out, n = cpu.clip_polygon(SQ4, [2, 5], [2, -1])
    assert n == 4
    np.testing.assert_allclose(out, [[2, 0], [4, 0], [4, 4], [2, 4]], rtol=1e-5)
    out, n = cpu.clip_polygon(UNIT, [100, 0], [100, 1])
    assert n == 4
    np.testing.assert_allclose(out, UNIT, rtol=1e-5)
    out, n = cpu.clip_polygon(UNIT + 1, [0, 0], [0, 1])
    assert n == 0 and out.shape == (0, 2)


def test_intersection_and_iou():
    np.testing.assert_allclose(cpu.polygon_intersection(SQ4, SQ4B), [[2, 2], [4, 2], [4, 4], [2, 4]], rtol=1e-5)
    assert cpu.polygon_intersection(UNIT, UNIT + 2).shape == (0, 2)
    assert np.isclose(cpu.polygon_iou(SQ4, SQ4B), 4 / 28, rtol=1e-5)
    assert cpu.polygon_iou(UNIT, UNIT) == pytest.approx(1.0)
    assert cpu.polygon_iou(UNIT, UNIT + 2) == pytest.approx(0.0)


def test_should_merge_strict_threshold():
    assert cpu.should_merge(SQ4, SQ4B, 0.1)
    assert not cpu.should_merge(SQ4, SQ4B, 0.2)
    assert not cpu.should_merge(UNIT, UNIT, 1.0)
    assert cpu.should_merge(UNIT, UNIT, 0.999)


def test_orientation_quirk():
    """a clockwise (negative shoelace) clip polygon gives IoU 0, even with itself."""
    rev = UNIT[::-1].copy()
    assert cpu.polygon_iou(UNIT, rev) == 0.0
    assert cpu.polygon_iou(rev, rev) == 0.0


def test_normalize_polygon_all_variants():
    for start in range(4):
        for var in (np.vstack([UNIT[(i + start) % 4] for i in range(4)]),
                    np.vstack([UNIT[(start - i) % 4] for i in range(4)])):
            np.testing.assert_allclose(cpu.normalize_polygon(UNIT, var), UNIT, rtol=1e-5)


def test_standard_nms_three_to_two():
    polys = [SQ4, SQ4 + 1, SQ4 + 10]
    kp, ks = cpu.standard_nms(polys, [0.9, 0.8, 0.7], 0.1)
    assert len(kp) == 2
    np.testing.assert_array_equal(ks, [0.9, 0.7])


def test_lanms_four_to_two_and_empty():
    boxes = np.array([[0, 0, 4, 0, 4, 4, 0, 4, 0.9], [1, 1, 5, 1, 5, 5, 1, 5, 0.8],
                      [10, 10, 14, 10, 14, 14, 10, 14, 0.7], [11, 11, 15, 11, 15, 15, 11, 15, 0.6]], np.float32)
    out = cpu.locality_aware_nms(boxes, 0.1)
    assert out.shape == (2, 9) and out.dtype == np.float32
    assert cpu.locality_aware_nms(np.zeros((0, 9), np.float32), 0.5).shape == (0, 9)
    assert cpu.locality_aware_nms(None, 0.5).shape == (0, 9)


def test_linear_resize_matches_cv2_when_shrinking_and_enlarging():
    """orc_resize_linear_u8c3 against cv2.resize(INTER_LINEAR) itself: the detector-input resize of EAST.predict
    (infer.py:304) shrinks whole pages, a case the ResizeAndPadA goldens (INTER_LINEAR only when enlarging) do not
    contain.  Third-party arithmetic: pinned to the OpenCV in this image (4.x)."""
    import cv2
    import numpy as np

    from oracle import cpu

    rng = np.random.default_rng(21)
    for (h, w), (dh, dw) in [((333, 517), (256, 256)), ((700, 900), (512, 512)), ((100, 180), (256, 256)),
                             ((512, 512), (256, 256)), ((1031, 777), (640, 640)), ((64, 2000), (128, 128)),
                             ((37, 41), (300, 7)), ((5, 5), (1, 1))]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        np.testing.assert_array_equal(cpu.cv_resize(img, (dw, dh), "linear"), cv2.resize(img, (dw, dh)),
                                      err_msg=f"{(h, w)} -> {(dh, dw)}")
