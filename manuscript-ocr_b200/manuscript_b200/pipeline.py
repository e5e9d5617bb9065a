"""Pipeline with the reference's duck-typed detector / recogniser contract (reference _pipeline.py:18-176).

detector.predict(image, vis=False, profile=...) -> {"page": Page, ...} | (Page, ...) | Page
recognizer.predict(List[np.ndarray RGB u8]) -> List[{"text", "confidence"}] (or (text, conf) tuples)

Between them: reading-order sort (host, as in the reference), the integer crop rectangles and -- when the
recogniser is this package's TRBA -- crop + resize-and-pad + normalise straight into the recogniser's
device batch in one kernel launch.  A foreign recogniser gets the list of uint8 crops it expects.
"""
import time

import numpy as np

from . import ops
from .east import read_image
from .reading_order import reorder_words
from .trba import TRBA


class Pipeline:
    def __init__(self, detector=None, recognizer=None, min_text_size=5):
        if detector is None or recognizer is None:
            raise ValueError("Pipeline(detector=..., recognizer=...) are required: the default EAST()/TRBA() of the "
                             "reference download network weights, which are outside this package")
        self.detector = detector
        self.recognizer = recognizer
        self.min_text_size = min_text_size

    def predict(self, image, recognize_text=True, vis=False, profile=False):
        start_time = time.time()
        t0 = time.time()
        det_out = self.detector.predict(image, vis=False, profile=profile)
        if isinstance(det_out, dict):
            page = det_out.get("page")
        elif isinstance(det_out, tuple):
            page = det_out[0]
        else:
            page = det_out
        if page is None:
            raise RuntimeError("Detector did not return a Page result.")
        if profile:
            print(f"Detection: {time.time() - t0:.3f}s")
        if vis:
            raise NotImplementedError("visualisation is outside the hot path: use the reference's visualize_page "
                                      "on the returned Page")
        if not recognize_text:
            return page

        image_array = read_image(image)
        img_h, img_w = image_array.shape[:2]
        t0 = time.time()
        all_words, all_rects = [], []
        for block in page.blocks:
            block.words = reorder_words(block.words)  # _pipeline.py:105-123
            if not block.words:
                continue
            polys = np.array([w.polygon for w in block.words], dtype=np.float32).reshape(len(block.words), -1)
            rects, valid = ops.word_rects(polys, img_h, img_w, self.min_text_size)  # _pipeline.py:125-137,204-221
            for w, r, ok in zip(block.words, rects, valid):
                if ok:
                    all_words.append(w)
                    all_rects.append(r)
        if profile:
            print(f"Extract {len(all_words)} crops: {time.time() - t0:.3f}s")

        if all_words:
            t0 = time.time()
            rects = np.stack(all_rects).astype(np.int32)
            if isinstance(self.recognizer, TRBA):
                results = []
                bs = self.recognizer.batch_size
                img3 = TRBA._as_rgb(image_array)
                for i in range(0, len(rects), bs):
                    batch = ops.crop_resize_pad(img3, rects[i:i + bs], self.recognizer.img_h, self.recognizer.img_w)
                    results.extend(self.recognizer.predict_batch(
                        self.recognizer.torch.from_numpy(batch).to(self.recognizer.device)))
            else:
                crops = [image_array[r[1]:r[3], r[0]:r[2]] for r in rects]
                results = self.recognizer.predict(crops)
            if profile:
                print(f"Recognition: {time.time() - t0:.3f}s")
            for word, result in zip(all_words, results):  # _pipeline.py:149-162
                if isinstance(result, dict):
                    text, confidence = result.get("text", ""), result.get("confidence", None)
                elif isinstance(result, tuple) and len(result) == 2:
                    text, confidence = result
                else:
                    text, confidence = (str(result) if result is not None else ""), None
                word.text = text
                word.recognition_confidence = confidence
        if profile:
            print(f"Pipeline total: {time.time() - start_time:.3f}s")
        return page

    def get_text(self, page):
        """_pipeline.py:193-202."""
        lines = []
        for block in page.blocks:
            sorted_words = sorted(block.words, key=lambda w: min(p[0] for p in w.polygon))
            texts = [w.text for w in sorted_words if getattr(w, "text", None)]
            if texts:
                lines.append(" ".join(texts))
        return "\n".join(lines)
