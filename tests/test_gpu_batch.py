"""The batched device path (ms_page_batch) against the oracle chain page by page, plus
size-independent properties at BASELINE.json's full shapes."""
import numpy as np
import pytest

import synthdata
from oracle import cpu

pytestmark = pytest.mark.gpu


def oracle_page(score, geo, img, page, out_hw=(32, 128), min_text=5):
    quads = cpu.decode_quads_from_maps(score, geo, 0.6, 4.0, 2)
    nms = cpu.locality_aware_nms(quads, 0.2)
    boxes = cpu.east_postprocess(nms, (page, page), target_size=page)
    rects, valid = cpu.word_rects(boxes, page, page, min_text)
    return quads, nms, boxes, rects[valid]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available()
    return torch


@pytest.mark.parametrize("page,words,n_pages", [(512, 80, 5), (1280, 500, 3)])
def test_page_batch_vs_oracle(torch_cuda, page, words, n_pages):
    torch = torch_cuda
    import manuscript_b200 as mb

    seeds = list(range(40, 40 + n_pages))
    score, geo, imgs = synthdata.make_batch(seeds, page, words)
    # one empty page in the middle (ragged batch)
    score[1] = 0.0
    params = mb.EastParams.default(target_size=page)
    runner = mb.PageBatch(device=0, params=params, cap_boxes=1024)
    res = runner.run(torch.from_numpy(score).cuda(), torch.from_numpy(geo).cuda(), torch.from_numpy(imgs).cuda())
    torch.cuda.synchronize()
    counts = res.box_counts.cpu().numpy()
    boxes = res.boxes.cpu().numpy()
    n_crops = int(res.n_crops.cpu()[0])
    crops = res.crops.cpu().numpy()[:n_crops]
    batch = res.batch[:n_crops].cpu().numpy()
    assert int(res.flags.cpu().numpy().max()) == 0
    k = 0
    for p in range(n_pages):
        _, _, want_boxes, want_rects = oracle_page(score[p], geo[p], imgs[p], page)
        assert counts[p] == len(want_boxes)
        np.testing.assert_array_equal(boxes[p, : counts[p]], want_boxes)
        mine = crops[crops[:, 0] == p]
        np.testing.assert_array_equal(mine[:, 1:], want_rects)
        for r in want_rects[:: max(1, len(want_rects) // 25)]:
            pass
        for j, r in enumerate(want_rects):
            if j % 7 == 0:
                _, chw = cpu.crop_resize_pad(imgs[p], r, 32, 128)
                np.testing.assert_array_equal(batch[k + j], chw)
        k += len(want_rects)
    assert k == n_crops and counts[1] == 0
    # the same batch through the host-buffer entry point
    res_h = runner.run_host(score, geo, imgs)
    np.testing.assert_array_equal(res_h.box_counts, counts)
    np.testing.assert_array_equal(res_h.boxes[0, : counts[0]], boxes[0, : counts[0]])
    assert int(res_h.n_crops[0]) == n_crops
    np.testing.assert_array_equal(res_h.crops[:n_crops], crops)
    assert res_h.batch.is_cuda and tuple(res_h.batch.shape) == (n_crops, 3, 32, 128)
    np.testing.assert_array_equal(res_h.batch.cpu().numpy(), batch)  # device-resident batch of the host entry point
    # pinned host tensors: the geometry is not uploaded, the decode kernel gathers it from host memory (zero copy)
    pin = [torch.from_numpy(a).pin_memory() for a in (score, geo, imgs)]
    res_p = runner.run_host(*pin)
    np.testing.assert_array_equal(res_p.box_counts, counts)
    for p in range(n_pages):
        np.testing.assert_array_equal(res_p.boxes[p, : counts[p]], boxes[p, : counts[p]])
    np.testing.assert_array_equal(res_p.crops[:n_crops], crops)
    np.testing.assert_array_equal(res_p.batch.cpu().numpy(), batch)
    # capacity knob of the NMS neighbour-pair buffer (ms_set_edge_factor / ms_get_edge_factor)
    assert runner.ctx.edge_factor == 16
    runner.ctx.edge_factor = 64
    assert runner.ctx.edge_factor == 64
    res_k = runner.run_host(score, geo, imgs)
    np.testing.assert_array_equal(res_k.box_counts, counts)
    runner.ctx.edge_factor = 16


def test_page_batch_original_size_pages(torch_cuda):
    """Maps of a 512-target detector, page images at their original (non-square, larger) size: the boxes are
    scaled to the original image (infer.py:134-147) before the filters and the crops are cut from it
    (_pipeline.py:125-137), as EAST.predict + Pipeline.predict do."""
    torch = torch_cuda
    import manuscript_b200 as mb

    target, oh, ow = 512, 1100, 830
    score, geo, _ = synthdata.make_batch([7, 8], target, 80)
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, (2, oh, ow, 3), dtype=np.uint8)
    runner = mb.PageBatch(device=0, params=mb.EastParams.default(target_size=target), cap_boxes=1024)
    res = runner.run_host(score, geo, imgs)
    n_crops = int(res.n_crops[0])
    batch = res.batch.cpu().numpy()
    k = 0
    for p in range(2):
        quads = cpu.decode_quads_from_maps(score[p], geo[p], 0.6, 4.0, 2)
        nms = cpu.locality_aware_nms(quads, 0.2)
        want = cpu.east_postprocess(nms, (oh, ow), target_size=target)
        rects, valid = cpu.word_rects(want, oh, ow, 5)
        rects = rects[valid]
        np.testing.assert_array_equal(res.page_boxes(p), want)
        np.testing.assert_array_equal(res.crops[k:k + len(rects), 1:], rects)
        for j in range(0, len(rects), 5):
            np.testing.assert_array_equal(batch[k + j], cpu.crop_resize_pad(imgs[p], rects[j], 32, 128)[1])
        k += len(rects)
    assert k == n_crops and k > 50


def test_full_size_properties(torch_cuda):
    """2048x2048 / ~2000 quads (BASELINE configs[2] shape), 2 pages: properties that need no oracle run --
    page-permutation equivariance, idempotence of NMS on its own output, kept rows are descending in
    score, every crop rectangle lies inside the page."""
    torch = torch_cuda
    import manuscript_b200 as mb

    page, words = 2048, 2000
    score, geo, imgs = synthdata.make_batch([0, 1], page, words)
    params = mb.EastParams.default(target_size=page)
    runner = mb.PageBatch(device=0, params=params, cap_boxes=4096)
    s, g, im = torch.from_numpy(score).cuda(), torch.from_numpy(geo).cuda(), torch.from_numpy(imgs).cuda()
    r1 = runner.run(s, g, im)
    torch.cuda.synchronize()
    c1, b1 = r1.box_counts.cpu().numpy().copy(), r1.boxes.cpu().numpy().copy()
    n1 = int(r1.n_crops.cpu()[0])
    crops1 = r1.crops.cpu().numpy()[:n1].copy()
    sum1 = r1.batch[:n1].double().sum(dim=(1, 2, 3)).cpu().numpy().copy()
    assert (c1 == words).all(), c1
    r2 = runner.run(s.flip(0).contiguous(), g.flip(0).contiguous(), im.flip(0).contiguous())
    torch.cuda.synchronize()
    c2, b2 = r2.box_counts.cpu().numpy(), r2.boxes.cpu().numpy()
    np.testing.assert_array_equal(c2, c1[::-1])
    np.testing.assert_array_equal(b2[1, : c2[1]], b1[0, : c1[0]])
    np.testing.assert_array_equal(b2[0, : c2[0]], b1[1, : c1[1]])
    sum2 = r2.batch[:n1].double().sum(dim=(1, 2, 3)).cpu().numpy()
    np.testing.assert_array_equal(np.sort(sum1), np.sort(sum2))  # checksum of per-crop checksums
    assert (crops1[:, 1] >= 0).all() and (crops1[:, 3] <= page).all() and (crops1[:, 2] >= 0).all()
    assert (crops1[:, 4] <= page).all() and (crops1[:, 3] > crops1[:, 1]).all()
    # NMS output: descending score; running NMS again on it removes nothing
    quads = mb.decode_quads_from_maps(score[0], geo[0], 0.6, 4.0, 2)
    nms = mb.locality_aware_nms(quads, 0.2)
    assert len(nms) == words and (np.diff(nms[:, 8]) <= 0).all()
    again = mb.locality_aware_nms(nms, 0.2)
    np.testing.assert_array_equal(rows(again), rows(nms))
    # and the page agrees with the oracle end to end on a bounded sample (decode exact, NMS on the oracle)
    np.testing.assert_array_equal(quads, cpu.decode_quads_from_maps(score[0], geo[0], 0.6, 4.0, 2))
    np.testing.assert_array_equal(nms, cpu.locality_aware_nms(quads, 0.2))


def rows(a):
    a = np.ascontiguousarray(a)
    return a[np.lexsort(a.T[::-1])]


def _sha(*arrs):
    import hashlib

    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def test_benchmarked_shape_vs_reference_golden_and_oracle(torch_cuda, golden_dir):
    """The shape bench.py measures (BASELINE configs[2]: 2048x2048 pages, ~2000 words, 32x128 crops), 3 pages of the
    bench's own corpus (seeds 0..2) through ms_page_batch -- so the 6000 crops go through the persistent TMA kernel
    exactly as they do in the benchmark:
      * page 0 against the REAL reference's outputs committed by tests/golden/make_golden_large.py (kept rows,
        final boxes, crop rectangles by sha256; every 40th crop's uint8 canvas from the reference's ResizeAndPadA);
      * every page: boxes and crop rectangles against the oracle, and EVERY 10th crop (600 crops) of the float32
        batch against the oracle's cpu.crop_resize_pad, bit for bit."""
    import os

    torch = torch_cuda
    import manuscript_b200 as mb

    g = np.load(os.path.join(golden_dir, "large_pages.npz"))
    page, words = 2048, 2000
    seeds = [0, 1, 2]
    score, geo, imgs = synthdata.make_batch(seeds, page, words)
    assert _sha(score[0], geo[0]) == str(g["cfg2_input_sha"]) and _sha(imgs[0]) == str(g["cfg2_image_sha"])
    runner = mb.PageBatch(device=0, params=mb.EastParams.default(target_size=page), cap_boxes=4096)
    res = runner.run(torch.from_numpy(score).cuda(), torch.from_numpy(geo).cuda(), torch.from_numpy(imgs).cuda())
    torch.cuda.synchronize()
    res.raise_for_flags()
    counts = res.box_counts.cpu().numpy()
    boxes = res.boxes.cpu().numpy()
    n_crops = int(res.n_crops.cpu()[0])
    crops = res.crops.cpu().numpy()[:n_crops]
    batch = res.batch[:n_crops].cpu().numpy()
    # page 0 vs the reference
    assert counts[0] == int(g["cfg2_n_final"]) and _sha(boxes[0, : counts[0]]) == str(g["cfg2_final_sha"])
    mine0 = crops[crops[:, 0] == 0][:, 1:]
    assert _sha(mine0.astype(np.int32)) == str(g["cfg2_rects_sha"])
    inv = np.float32(1 / 127.5)
    for k, canvas in enumerate(g["cfg2_canvas_every40"]):
        want = ((canvas.astype(np.float32) - np.float32(127.5)) * inv).transpose(2, 0, 1)
        np.testing.assert_array_equal(batch[40 * k], want)
    nms0 = mb.locality_aware_nms(mb.decode_quads_from_maps(score[0], geo[0], 0.6, 4.0, 2), 0.2)
    np.testing.assert_array_equal(nms0, g["cfg2_lanms_stable"])
    # every page vs the oracle
    k = checked = 0
    for p in range(len(seeds)):
        _, _, want_boxes, want_rects = oracle_page(score[p], geo[p], imgs[p], page)
        np.testing.assert_array_equal(boxes[p, : counts[p]], want_boxes)
        mine = crops[crops[:, 0] == p]
        np.testing.assert_array_equal(mine[:, 1:], want_rects)
        for j in range(0, len(want_rects), 10):
            np.testing.assert_array_equal(batch[k + j], cpu.crop_resize_pad(imgs[p], want_rects[j], 32, 128)[1])
            checked += 1
        k += len(want_rects)
    assert k == n_crops and checked >= 600


def test_stress_4096_page(torch_cuda, golden_dir):
    """BASELINE configs[3]: one 4096x4096 page, ~10k quads, ~75k candidates (NMS-bound).  Decode is compared with
    the oracle bit for bit; NMS through its properties (the O(n^2) oracle would take minutes): one box per word,
    descending scores, idempotent, every kept box overlaps exactly one ground-truth word."""
    torch = torch_cuda
    import manuscript_b200 as mb

    page, words = 4096, 10000
    score, geo, gt = synthdata.make_maps(3, page, words)
    quads = mb.decode_quads_from_maps(score, geo, 0.6, 4.0, 2)
    np.testing.assert_array_equal(quads, cpu.decode_quads_from_maps(score, geo, 0.6, 4.0, 2))
    assert len(quads) > 50000
    nms = mb.locality_aware_nms(quads, 0.2)
    assert len(nms) == words and (np.diff(nms[:, 8]) <= 0).all()
    np.testing.assert_array_equal(rows(mb.locality_aware_nms(nms, 0.2)), rows(nms))
    # each kept quad sits on its own word: nearest ground-truth centre is unique
    cen = nms[:, :8].reshape(-1, 4, 2).mean(axis=1)
    gcen = gt.mean(axis=1)
    order = np.argsort(gcen[:, 1] * 8192 + gcen[:, 0])
    # brute force in chunks
    owner = np.empty(len(cen), np.int64)
    for i in range(0, len(cen), 1000):
        d = ((cen[i:i + 1000, None, :] - gcen[None, :, :]) ** 2).sum(-1)
        owner[i:i + 1000] = d.argmin(1)
    assert len(np.unique(owner)) == words
    # the same page through the batched path
    imgs = synthdata.make_page_image(3, page)[None]
    runner = mb.PageBatch(device=0, params=mb.EastParams.default(target_size=page), cap_boxes=16384)
    res = runner.run(torch.from_numpy(score[None]).cuda(), torch.from_numpy(geo[None]).cuda(),
                     torch.from_numpy(imgs).cuda())
    torch.cuda.synchronize()
    assert int(res.flags.cpu()[0]) == 0 and int(res.box_counts.cpu()[0]) == words
    want = cpu.east_postprocess(nms, (page, page), target_size=page)
    np.testing.assert_array_equal(res.boxes[0, :words].cpu().numpy(), want)
    del order
    # ... and against the digests of the offline oracle run (tests/golden/make_golden_large.py; the oracle needs
    # ~40 s for this page, so its outputs are committed as sha256 instead of being recomputed here)
    import os

    g = np.load(os.path.join(golden_dir, "large_pages.npz"))
    assert _sha(score, geo) == str(g["cfg3_input_sha"])
    assert len(quads) == int(g["cfg3_n_candidates"]) and _sha(quads) == str(g["cfg3_quads_sha"])
    assert len(nms) == int(g["cfg3_n_kept"]) and _sha(nms) == str(g["cfg3_lanms_sha"])
    assert _sha(res.boxes[0, :words].cpu().numpy()) == str(g["cfg3_final_sha"])


def test_host_entry_chunks_reading_order_and_odd_shapes(torch_cuda):
    """ms_page_batch_host cuts the batch into chunks (H2D of chunk c+1 under the kernels of chunk c), uploads only the
    geometry rows a quantised decode reads, and appends every chunk's crops: same results as the one-shot device
    path, with the reading-order stage on, for 11 pages (chunks of 2) and for quantisation 1 / 4."""
    torch = torch_cuda
    import manuscript_b200 as mb

    page, words, n_pages = 512, 60, 11
    score, geo, imgs = synthdata.make_batch(list(range(70, 70 + n_pages)), page, words)
    score[4] = 0.0  # an empty page inside a chunk
    for q, ro in [(2, 1), (1, 0), (4, 1)]:
        params = mb.EastParams.default(target_size=page, quantization=q, sort_reading_order=ro)
        runner = mb.PageBatch(device=0, params=params, cap_boxes=512)
        res = runner.run(torch.from_numpy(score).cuda(), torch.from_numpy(geo).cuda(), torch.from_numpy(imgs).cuda())
        torch.cuda.synchronize()
        counts = res.box_counts.cpu().numpy().copy()
        boxes = res.boxes.cpu().numpy().copy()
        n_crops = int(res.n_crops.cpu()[0])
        crops = res.crops.cpu().numpy()[:n_crops].copy()
        batch = res.batch[:n_crops].cpu().numpy().copy()
        assert int(res.flags.cpu().numpy().max()) == 0 and counts[4] == 0 and n_crops > 0
        rh = runner.run_host(score, geo, imgs)
        np.testing.assert_array_equal(rh.box_counts, counts)
        for p in range(n_pages):
            np.testing.assert_array_equal(rh.boxes[p, : counts[p]], boxes[p, : counts[p]])
        assert int(rh.n_crops[0]) == n_crops
        np.testing.assert_array_equal(rh.crops[:n_crops], crops)
        np.testing.assert_array_equal(rh.batch.cpu().numpy(), batch)
        # page 0 against the oracle, in the order the stage promises
        quads = cpu.decode_quads_from_maps(score[0], geo[0], 0.6, 4.0, q)
        want = cpu.east_postprocess(cpu.locality_aware_nms(quads, 0.2), (page, page), target_size=page)
        if ro:
            want = want[mb.word_reading_order(want[:, :8])]
        np.testing.assert_array_equal(boxes[0, : counts[0]], want)


def test_boxes_only_and_capacity_flags(torch_cuda):
    """Without page images the batch stops after the box filters; a page with more boxes than cap_boxes is flagged
    (device path) / raises (host path) instead of being silently truncated."""
    torch = torch_cuda
    import manuscript_b200 as mb

    page, words = 512, 80
    score, geo, imgs = synthdata.make_batch([90, 91], page, words)
    params = mb.EastParams.default(target_size=page)
    runner = mb.PageBatch(device=0, params=params, cap_boxes=256, want_batch=False)
    res = runner.run(torch.from_numpy(score).cuda(), torch.from_numpy(geo).cuda(), None)
    torch.cuda.synchronize()
    counts = res.box_counts.cpu().numpy()
    assert (counts == words).all() and int(res.flags.cpu().numpy().max()) == 0 and res.batch is None
    want = oracle_page(score[1], geo[1], imgs[1], page)[2]
    np.testing.assert_array_equal(res.boxes[1, : counts[1]].cpu().numpy(), want)

    small = mb.PageBatch(device=0, params=params, cap_boxes=32)
    r2 = small.run(torch.from_numpy(score).cuda(), torch.from_numpy(geo).cuda(), torch.from_numpy(imgs).cuda())
    torch.cuda.synchronize()
    assert (r2.flags.cpu().numpy() & 1).all() and (r2.box_counts.cpu().numpy() == 32).all()
    with pytest.raises(mb.CABIError) as e:
        small.run_host(score, geo, imgs)
    assert e.value.code == -3
    # odd map size with quantisation 2: the reference raises IndexError (utils.py:370)
    bad_score = np.zeros((1, 7, 8), np.float32)
    bad_score[0, 6, 0] = 0.9
    with pytest.raises(IndexError):
        mb.PageBatch(device=0, params=mb.EastParams.default(target_size=32), cap_boxes=8, want_batch=False).run_host(
            bad_score, np.zeros((1, 8, 7, 8), np.float32), None)


def test_page_batch_ragged_originals(torch_cuda):
    """A batch of original images of different sizes (one narrower than 16 pixels' alignment, one with an odd width so
    that its rows start at every misalignment) with maps of one 512-target detector: per page, boxes scaled to that
    page's size, crops cut from its pixels -- through the device entry (list of CUDA tensors) and the host entry."""
    torch = torch_cuda
    import manuscript_b200 as mb

    target = 512
    sizes = [(700, 900), (512, 512), (333, 1001), (1200, 640)]
    seeds = [21, 22, 23, 24]
    score, geo, _ = synthdata.make_batch(seeds, target, 70)
    rng = np.random.default_rng(4)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    runner = mb.PageBatch(device=0, params=mb.EastParams.default(target_size=target), cap_boxes=1024)
    res_d = runner.run_ragged(torch.from_numpy(score).cuda(), torch.from_numpy(geo).cuda(),
                              [torch.from_numpy(im).cuda() for im in imgs])
    torch.cuda.synchronize()
    res_h = runner.run_host_ragged(score, geo, imgs)
    n = int(res_h.n_crops[0])
    assert int(res_d.n_crops.cpu()[0]) == n
    np.testing.assert_array_equal(res_d.crops[:n].cpu().numpy(), res_h.crops[:n])
    np.testing.assert_array_equal(res_d.batch[:n].cpu().numpy(), res_h.batch.cpu().numpy())
    batch = res_h.batch.cpu().numpy()
    k = 0
    for p, (oh, ow) in enumerate(sizes):
        nms = cpu.locality_aware_nms(cpu.decode_quads_from_maps(score[p], geo[p], 0.6, 4.0, 2), 0.2)
        want = cpu.east_postprocess(nms, (oh, ow), target_size=target)
        rects, valid = cpu.word_rects(want, oh, ow, 5)
        rects = rects[valid]
        np.testing.assert_array_equal(res_h.page_boxes(p), want)
        np.testing.assert_array_equal(res_d.boxes[p, : len(want)].cpu().numpy(), want)
        rows = res_h.crops[k:k + len(rects)]
        assert (rows[:, 0] == p).all()
        np.testing.assert_array_equal(rows[:, 1:], rects)
        for j in range(0, len(rects), 3):
            np.testing.assert_array_equal(batch[k + j], cpu.crop_resize_pad(imgs[p], rects[j], 32, 128)[1])
        k += len(rects)
    assert k == n and n > 150


def test_repeated_batches_replay_a_cuda_graph(torch_cuda):
    """A repeated ms_page_batch call (same buffers and sizes) is captured into a CUDA graph at its second occurrence
    and replayed afterwards: same results, same launch accounting, and new data in the same buffers is honoured."""
    torch = torch_cuda
    import manuscript_b200 as mb

    page = 512
    score, geo, imgs = synthdata.make_batch([61, 62], page, 60)
    score2, geo2, imgs2 = synthdata.make_batch([63, 64], page, 45)
    runner = mb.PageBatch(device=0, params=mb.EastParams.default(target_size=page), cap_boxes=512)
    s, g, im = torch.from_numpy(score).cuda(), torch.from_numpy(geo).cuda(), torch.from_numpy(imgs).cuda()

    def snapshot():
        r = runner.run(s, g, im)
        torch.cuda.synchronize()
        n = int(r.n_crops.cpu()[0])
        return (r.box_counts.cpu().numpy().copy(), r.boxes.cpu().numpy().copy(), r.crops[:n].cpu().numpy().copy(),
                r.batch[:n].cpu().numpy().copy())

    l0 = runner.ctx.launches
    first = snapshot()       # direct launches
    per_call = runner.ctx.launches - l0
    second = snapshot()      # captured + replayed
    third = snapshot()       # replayed
    assert runner.ctx.launches - l0 == 3 * per_call and per_call > 20
    for a, b, c in zip(first, second, third):
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(a, c)
    # new pages written into the same device buffers
    s.copy_(torch.from_numpy(score2)), g.copy_(torch.from_numpy(geo2)), im.copy_(torch.from_numpy(imgs2))
    fourth = snapshot()
    fresh = mb.PageBatch(device=0, params=mb.EastParams.default(target_size=page), cap_boxes=512)
    r = fresh.run(s.clone(), g.clone(), im.clone())
    torch.cuda.synchronize()
    n = int(r.n_crops.cpu()[0])
    np.testing.assert_array_equal(fourth[0], r.box_counts.cpu().numpy())
    np.testing.assert_array_equal(fourth[2], r.crops[:n].cpu().numpy())
    np.testing.assert_array_equal(fourth[3], r.batch[:n].cpu().numpy())
    assert not np.array_equal(first[0], fourth[0])
    _, _, want_boxes, _ = oracle_page(score2[1], geo2[1], imgs2[1], page)
    np.testing.assert_array_equal(fourth[1][1, : fourth[0][1]], want_boxes)


def test_front_stages_as_two_concurrent_halves(torch_cuda):
    """A batch of 16 pages or more runs its front stages (decode -> LANMS -> filters -> reading order) as two halves on
    two streams.  Same bits as one sequence (MS_B200_NO_SPLIT=1), as a graph replay too; pages of both halves against
    the oracle; an odd page count, an empty page and the reading-order stage included."""
    import os

    torch = torch_cuda
    import manuscript_b200 as mb

    page, words, n_pages = 512, 80, 19
    score, geo, imgs = synthdata.make_batch(list(range(300, 300 + n_pages)), page, words)
    score[11] = 0.0
    d = [torch.from_numpy(x).cuda() for x in (score, geo, imgs)]
    for ro in (0, 1):
        params = mb.EastParams.default(target_size=page, sort_reading_order=ro)
        split = mb.PageBatch(device=0, params=params, cap_boxes=512)
        os.environ["MS_B200_NO_SPLIT"] = "1"
        try:
            plain = mb.PageBatch(device=0, params=params, cap_boxes=512)  # its context is created here
        finally:
            del os.environ["MS_B200_NO_SPLIT"]
        want = plain.run(*d)
        torch.cuda.synchronize()
        wb, wc, wf = want.boxes.cpu().numpy().copy(), want.box_counts.cpu().numpy().copy(), want.flags.cpu().numpy().copy()
        wn = int(want.n_crops.cpu()[0])
        wcr, wba = want.crops.cpu().numpy()[:wn].copy(), want.batch[:wn].cpu().numpy().copy()
        for rep in range(3):  # direct launches, graph capture, graph replay
            got = split.run(*d)
            torch.cuda.synchronize()
            np.testing.assert_array_equal(got.box_counts.cpu().numpy(), wc)
            np.testing.assert_array_equal(got.flags.cpu().numpy(), wf)
            for pg in range(n_pages):
                np.testing.assert_array_equal(got.boxes[pg, : wc[pg]].cpu().numpy(), wb[pg, : wc[pg]])
            assert int(got.n_crops.cpu()[0]) == wn
            np.testing.assert_array_equal(got.crops.cpu().numpy()[:wn], wcr)
            np.testing.assert_array_equal(got.batch[:wn].cpu().numpy(), wba)
        assert int(wf.max()) == 0 and wc[11] == 0
        if ro == 0:
            for pg in (0, 8, 9, 18):  # both halves (9 + 10 pages)
                _, _, want_boxes, want_rects = oracle_page(score[pg], geo[pg], imgs[pg], page)
                np.testing.assert_array_equal(wb[pg, : wc[pg]], want_boxes)
                np.testing.assert_array_equal(wcr[wcr[:, 0] == pg][:, 1:], want_rects)


_PDL_CHILD = r"""
import hashlib, os, sys
sys.path[:0] = [ROOT, os.path.join(ROOT, "manuscript-ocr_b200")]
import numpy as np, torch
import manuscript_b200 as mb, synthdata
page, n_pages = 512, 18
score, geo, imgs = synthdata.make_batch(list(range(700, 700 + n_pages)), page, 70)
d = [torch.from_numpy(x).cuda() for x in (score, geo, imgs)]
h = hashlib.sha256()
for ro in (0, 1):
    runner = mb.PageBatch(device=0, params=mb.EastParams.default(target_size=page, sort_reading_order=ro), cap_boxes=512)
    for rep in range(3):  # direct launches, graph capture, graph replay
        r = runner.run(*d)
        torch.cuda.synchronize()
        n = int(r.n_crops.cpu()[0])
        cnt = r.box_counts.cpu().numpy()
        h.update(cnt.tobytes()); h.update(r.flags.cpu().numpy().tobytes())
        for pg in range(n_pages):
            h.update(r.boxes[pg, : cnt[pg]].cpu().numpy().tobytes())
        h.update(r.crops[:n].cpu().numpy().tobytes()); h.update(r.batch[:n].cpu().numpy().tobytes())
print("DIGEST", h.hexdigest(), int(cnt.sum()), n)
"""


def test_programmatic_dependent_launch_changes_nothing(torch_cuda):
    """Every kernel is launched as a programmatic dependent launch and waits (griddepcontrol.wait) before it touches
    memory.  The switch is read once per process, so the same batch -- two concurrent halves, with and without the
    reading-order stage, direct launches and graph replay -- runs in two child processes, with MS_B200_NO_PDL=1 and
    without, and must give the same bits."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {}
    for tag, extra in (("pdl", {}), ("plain", {"MS_B200_NO_PDL": "1"})):
        env = dict(os.environ, **extra)
        env.pop("MS_B200_NO_PDL", None) if not extra else None
        p = subprocess.run([sys.executable, "-c", f"ROOT = {root!r}\n" + _PDL_CHILD], env=env, capture_output=True,
                           text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("DIGEST")][-1].split()
        out[tag] = line[1:]
    assert out["pdl"] == out["plain"], out
    assert int(out["pdl"][1]) > 500 and int(out["pdl"][2]) > 500
