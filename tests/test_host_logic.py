"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol the header declares, the host
logic (reading order, sharding, argument checks) matches the reference / its contract, the product never
imports the oracle, and compute calls fail loudly without a device."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "manuscript-ocr_b200")
HEADER = os.path.join(ROOT, "include", "manuscript_b200.h")


def header_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"^MS_API\s+[\w \*]+?\b(ms_\w+)\s*\(", src, flags=re.M)))


def test_library_exports_every_header_symbol():
    import manuscript_b200 as mb
    from manuscript_b200 import _cabi

    syms = header_symbols()
    assert len(syms) >= 24, syms
    assert sorted(_cabi.SIGNATURES) == syms  # the ctypes table and the header agree
    lib = mb.load_library()  # types every entry; AttributeError on a missing export
    for s in syms:
        assert getattr(lib, s) is not None
    assert b"sm_100a" in lib.ms_version()
    p = mb.EastParams.default()
    assert (p.quantization, p.target_size, p.anomaly_min_box_count) == (2, 1280, 30)
    assert abs(p.score_thresh - 0.6) < 1e-6 and p.scale == 4.0 and p.iou_threshold == 0.2


def test_no_device_is_loud():
    import torch

    import manuscript_b200 as mb

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mb.CABIError) as e:
        mb.Context(0)
    assert e.value.code == -5
    with pytest.raises(mb.CABIError):
        mb.decode_quads_from_maps(np.zeros((4, 4), np.float32), np.zeros((8, 4, 4), np.float32), 0.5, 4.0)
    with pytest.raises(RuntimeError):
        mb.PageBatch(device=0)
    with pytest.raises(RuntimeError):
        mb.EAST()


def test_product_never_touches_the_oracle():
    bad = []
    for d, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "liboracle" in txt:
                    bad.append(f)
    assert bad == []


def test_reading_order_golden(golden_dir):
    import manuscript_b200 as mb

    g = np.load(os.path.join(golden_dir, "reading_order.npz"))
    n = int(g["n_cases"])
    assert n >= 20
    for i in range(n):
        boxes = [tuple(int(v) for v in r) for r in g[f"boxes_{i}"]]
        got = np.array(mb.resolve_intersections(boxes), np.int64).reshape(-1, 4)
        np.testing.assert_array_equal(got, g[f"resolved_{i}"], err_msg=f"case {i}")
        got = np.array(mb.sort_boxes_reading_order(boxes), np.int64).reshape(-1, 4)
        np.testing.assert_array_equal(got, g[f"sorted_{i}"], err_msg=f"case {i}")
        got = np.array(mb.sort_boxes_reading_order_with_resolutions(boxes), np.int64).reshape(-1, 4)
        np.testing.assert_array_equal(got, g[f"sorted_res_{i}"], err_msg=f"case {i}")


def test_reorder_words_follows_pipeline_loop():
    """_pipeline.py:113-123: each sorted box picks the first word with the same integer bbox."""
    from manuscript_b200.reading_order import reorder_words
    from manuscript_b200.types import Word

    def word(x0, y0, x1, y1, tag):
        return Word(polygon=[(x0, y0), (x1, y0), (x1, y1), (x0, y1)], detection_confidence=0.9, text=tag)

    words = [word(300.7, 10.2, 380.1, 40.9, "c"), word(10.5, 12.0, 90.0, 41.0, "a"), word(100.2, 11.0, 180.0, 40.0, "b"),
             word(10.1, 80.0, 120.9, 110.0, "d")]
    assert [w.text for w in reorder_words(words, on_device=False)] == ["a", "b", "c", "d"]
    assert reorder_words([]) == []


def test_shard_pages_partitions():
    from manuscript_b200 import shard_pages

    for n in (0, 1, 7, 64, 1024, 1000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                rg = shard_pages(n, world, r)
                seen.extend(rg)
                assert len(rg) in (n // world, n // world + 1) or n < world
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        shard_pages(10, 2, 2)


def test_types_validate_like_the_reference():
    from manuscript_b200.types import Block, Page, Word

    w = Word(polygon=[(0, 0), (1, 0), (1, 1), (0, 1)], detection_confidence=0.5)
    assert w.text is None and w.recognition_confidence is None
    with pytest.raises(Exception):
        Word(polygon=[(0, 0)], detection_confidence=1.5)  # _types.py:9-11: 0 <= confidence <= 1
    assert Page(blocks=[Block(words=[w])]).blocks[0].words[0].polygon[2] == (1.0, 1.0)
