"""Pipeline with the reference's duck-typed detector / recogniser contract (reference _pipeline.py:18-176).

    detector.predict(image, vis=False, profile=...)  ->  {"page": Page, ...} | (Page, ...) | Page
    recognizer.predict(List[np.ndarray RGB u8])      ->  List[{"text", "confidence"}]  (or (text, conf) tuples)

Between the two calls sit the reading-order sort, the integer crop rectangles and -- when the recogniser is this
package's TRBA -- crop + resize-and-pad + normalise straight into the recogniser's device batch, one kernel launch per
recogniser batch.  A foreign recogniser gets the list of uint8 crops it expects (plain slices of the page image).
"""
import time

import numpy as np

from . import ops
from .east import read_image
from .reading_order import reorder_words
from .trba import TRBA


def _page_of(det_out):
    """_pipeline.py:69-77: the three shapes a detector may answer in."""
    if isinstance(det_out, dict):
        page = det_out.get("page")
    elif isinstance(det_out, tuple):
        page = det_out[0]
    else:
        page = det_out
    if page is None:
        raise RuntimeError("Detector did not return a Page result.")
    return page


def _text_and_confidence(result):
    """_pipeline.py:152-159: dict, (text, confidence) tuple or anything printable."""
    if isinstance(result, dict):
        return result.get("text", ""), result.get("confidence", None)
    if isinstance(result, tuple) and len(result) == 2:
        return result
    return (str(result) if result is not None else ""), None


class Pipeline:
    def __init__(self, detector=None, recognizer=None, min_text_size=5, rotated_crops=False):
        if detector is None or recognizer is None:
            raise ValueError("Pipeline(detector=..., recognizer=...) are required: the default EAST()/TRBA() of the "
                             "reference download network weights, which are outside this package")
        self.detector = detector
        self.recognizer = recognizer
        self.min_text_size = min_text_size
        # EXTENSION (reference todo.md:1): rectify every word quad with a perspective warp instead of
        # cutting its bounding rectangle.  Off by default = the reference's behaviour.
        self.rotated_crops = bool(rotated_crops)

    # ---- the steps between detector and recogniser -------------------------------------------------------------
    def _ordered_crop_rects(self, page, img_h, img_w):
        """Reading order per block (_pipeline.py:105-123, mutates block.words like the reference), then the crop
        rectangle of every word that passes the size filter (_pipeline.py:125-137, 204-221)."""
        words, rects = [], []
        for block in page.blocks:
            block.words = reorder_words(block.words)
            if not block.words:
                continue
            polys = np.array([w.polygon for w in block.words], dtype=np.float32).reshape(len(block.words), -1)
            block_rects, valid = ops.word_rects(polys, img_h, img_w, self.min_text_size)
            for word, rect, ok in zip(block.words, block_rects, valid):
                if ok:
                    words.append(word)
                    rects.append(rect)
        return words, (np.stack(rects).astype(np.int32) if rects else np.zeros((0, 4), np.int32))

    def _recognise_rotated(self, image_array, words):
        """rotated_crops=True: one rectified patch per word quad; words without a patch (smaller than min_text_size on
        a side, degenerate) are skipped like the reference skips small boxes.  Returns (kept words, results)."""
        rec = self.recognizer
        rgb = TRBA._as_rgb(image_array)
        quads = np.array([w.polygon for w in words], dtype=np.float32).reshape(len(words), -1)[:, :8]
        if not isinstance(rec, TRBA):
            patches = [ops.warp_quad(rgb, q) for q in quads]
            keep = [i for i, p in enumerate(patches)
                    if p is not None and min(p.shape[:2]) >= self.min_text_size]
            return [words[i] for i in keep], rec.predict([patches[i] for i in keep])
        results, kept = [], []
        for i in range(0, len(words), rec.batch_size):
            batch, valid = ops.quad_crop_resize_pad(rgb, quads[i:i + rec.batch_size], rec.img_h, rec.img_w,
                                                    self.min_text_size)
            if valid.any():
                results.extend(rec.predict_batch(rec.torch.from_numpy(batch[valid]).to(rec.device)))
                kept.extend(w for w, ok in zip(words[i:i + rec.batch_size], valid) if ok)
        return kept, results

    def _recognise(self, image_array, rects):
        rec = self.recognizer
        if not isinstance(rec, TRBA):
            return rec.predict([image_array[y1:y2, x1:x2] for x1, y1, x2, y2 in rects])
        results = []
        rgb = TRBA._as_rgb(image_array)
        for i in range(0, len(rects), rec.batch_size):
            batch = ops.crop_resize_pad(rgb, rects[i:i + rec.batch_size], rec.img_h, rec.img_w)
            results.extend(rec.predict_batch(rec.torch.from_numpy(batch).to(rec.device)))
        return results

    # ---- the reference's public surface ---------------------------------------------------------------------------
    def predict(self, image, recognize_text=True, vis=False, profile=False):
        t_start = t0 = time.time()
        page = _page_of(self.detector.predict(image, vis=False, profile=profile))
        if profile:
            print(f"Detection: {time.time() - t0:.3f}s")
        if vis:
            raise NotImplementedError("visualisation is outside the hot path: use the reference's visualize_page "
                                      "on the returned Page")
        if not recognize_text:
            return page

        image_array = read_image(image)
        t0 = time.time()
        words, rects = self._ordered_crop_rects(page, *image_array.shape[:2])
        if profile:
            print(f"Extract {len(words)} crops: {time.time() - t0:.3f}s")
        if words:
            t0 = time.time()
            if self.rotated_crops:
                words, results = self._recognise_rotated(image_array, [w for b in page.blocks for w in b.words])
            else:
                results = self._recognise(image_array, rects)
            if profile:
                print(f"Recognition: {time.time() - t0:.3f}s")
            for word, result in zip(words, results):
                word.text, word.recognition_confidence = _text_and_confidence(result)
        if profile:
            print(f"Pipeline total: {time.time() - t_start:.3f}s")
        return page

    def get_text(self, page):
        """_pipeline.py:193-202: one line per block, words left to right."""
        lines = []
        for block in page.blocks:
            texts = [w.text for w in sorted(block.words, key=lambda w: min(pt[0] for pt in w.polygon))
                     if getattr(w, "text", None)]
            if texts:
                lines.append(" ".join(texts))
        return "\n".join(lines)
