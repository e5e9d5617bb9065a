"""EAST detector wrapper with the reference's predict() contract, post-processing on the B200.

Mirrors `EAST` (reference detectors/_east/infer.py:28-132 constructor kwargs, 235-402 predict): image ->
resize -> network -> score / geometry maps -> decode -> LANMS -> expand -> scale -> box filters -> Page.
The network (a ResNet-FPN, reference detectors/_east/east.py) is outside the hot path and is supplied by the
caller as any callable / nn.Module returning {"score": (1,1,H,W), "geometry": (1,8,H,W)}; everything after it runs
in the CUDA kernels of this package and the maps never leave the device (the reference's
`.cpu().numpy()` at infer.py:312-313 is gone).
"""
import ctypes as C
import time
from pathlib import Path

import numpy as np

from ._cabi import MS_FLAG_EDGE_OVERFLOW, EastParams, check
from .batch import PageBatch, _raise_for_flags
from .reading_order import reorder_words
from .types import Block, Page, Word


def read_image(img_or_path):
    """reference detectors/_east/utils.py:477-497: path -> RGB uint8 array, ndarray passes through."""
    if isinstance(img_or_path, (str, Path)):
        import cv2

        img = cv2.imread(str(img_or_path))
        if img is None:
            try:
                from PIL import Image

                with Image.open(str(img_or_path)) as pil_img:
                    img = np.array(pil_img.convert("RGB"))
            except Exception as e:
                raise FileNotFoundError(f"Cannot read image with cv2 or PIL: {img_or_path}. Error: {e}")
        else:
            img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    elif isinstance(img_or_path, np.ndarray):
        img = img_or_path
    else:
        raise TypeError(f"Unsupported type for image input: {type(img_or_path)}")
    return img


class EAST:
    def __init__(self, model=None, target_size=1280, expand_ratio_w=0.9, expand_ratio_h=0.9, score_thresh=0.6,
                 iou_threshold=0.2, score_geo_scale=0.25, quantization=2, axis_aligned_output=True,
                 remove_area_anomalies=True, anomaly_sigma_threshold=5.0, anomaly_min_box_count=30, device=0,
                 cap_boxes=8192):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("manuscript_b200.EAST needs a CUDA device: there is no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", int(device)) if not isinstance(device, torch.device) else device
        self.model = model
        self.target_size = int(target_size)
        self.score_thresh = score_thresh
        self.score_geo_scale = score_geo_scale
        self.iou_threshold = iou_threshold
        self.expand_ratio_w = expand_ratio_w
        self.expand_ratio_h = expand_ratio_h
        self.quantization = quantization
        self.axis_aligned_output = axis_aligned_output
        self.remove_area_anomalies = remove_area_anomalies
        self.anomaly_sigma_threshold = anomaly_sigma_threshold
        self.anomaly_min_box_count = anomaly_min_box_count
        self._runner = PageBatch(device=self.device.index or 0, params=self._params(), cap_boxes=cap_boxes,
                                 want_batch=False)

    def _params(self):
        return EastParams.default(
            score_thresh=float(np.float32(self.score_thresh)), scale=1.0 / self.score_geo_scale,
            quantization=int(self.quantization), iou_threshold=float(self.iou_threshold),
            expand_ratio_w=float(self.expand_ratio_w), expand_ratio_h=float(self.expand_ratio_h),
            target_size=int(self.target_size), axis_aligned_output=int(bool(self.axis_aligned_output)),
            remove_area_anomalies=int(bool(self.remove_area_anomalies)),
            anomaly_sigma_threshold=float(self.anomaly_sigma_threshold),
            anomaly_min_box_count=int(self.anomaly_min_box_count))

    # ---- the hot path: maps (device or host) -> final boxes (K,9) float32 on the host -----------------------------
    def boxes_from_maps(self, score_map, geo_map, orig_hw):
        """score (H,W) / (1,H,W), geo (8,H,W) torch CUDA tensors or numpy arrays; orig_hw = (h, w) of the image
        the boxes are scaled back to (infer.py:346-348).  Runs infer.py:319-356 on the device."""
        torch = self.torch
        s = torch.as_tensor(score_map, dtype=torch.float32, device=self.device)
        g = torch.as_tensor(geo_map, dtype=torch.float32, device=self.device)
        if s.dim() == 3:
            s = s[0]
        H, W = s.shape
        if g.shape != (8, H, W):
            raise ValueError(f"geo_map must be (8,{H},{W}), got {tuple(g.shape)}")
        s = s.contiguous()
        g = g.contiguous()
        r = self._runner
        r.params = self._params()
        ctx, lib = r.ctx, r.ctx.lib
        cap = ((H + max(self.quantization, 1) - 1) // max(self.quantization, 1)) * \
              ((W + max(self.quantization, 1) - 1) // max(self.quantization, 1))
        cap = max(cap, 1)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        cand = torch.empty((cap, 9), dtype=torch.float32, device=self.device)
        kept = torch.empty((cap, 9), dtype=torch.float32, device=self.device)
        out = torch.empty((cap, 9), dtype=torch.float32, device=self.device)
        meta = torch.zeros((8,), dtype=torch.int32, device=self.device)  # n_cand, n_kept, n_out, flags, orig h, w
        meta[4] = int(orig_hw[0])
        meta[5] = int(orig_hw[1])
        p = r.params
        while True:
            meta[:4] = 0
            with torch.cuda.device(self.device):
                check(lib.ms_decode_quads(ctx.handle, s.data_ptr(), g.data_ptr(), 1, H, W, p.score_thresh, p.scale,
                                          p.quantization, cand.data_ptr(), cap, meta[0:].data_ptr(),
                                          meta[3:].data_ptr(), C.c_void_p(stream)))
                check(lib.ms_lanms(ctx.handle, cand.data_ptr(), meta[0:].data_ptr(), 1, cap, p.iou_threshold,
                                   kept.data_ptr(), meta[1:].data_ptr(), meta[3:].data_ptr(), C.c_void_p(stream)))
                check(lib.ms_east_boxes(ctx.handle, kept.data_ptr(), meta[1:].data_ptr(), 1, cap, C.byref(p),
                                        meta[4:].data_ptr(), out.data_ptr(), meta[2:].data_ptr(), C.c_void_p(stream)))
            m = meta.cpu().numpy()
            # very dense candidates: grow the NMS neighbour-pair capacity and run again
            if (int(m[3]) & MS_FLAG_EDGE_OVERFLOW) and not (int(m[3]) & 3) and ctx.grow_edge_factor():
                continue
            break
        _raise_for_flags(m[3:4])
        return out[: int(m[2])].cpu().numpy()

    def _forward(self, img):
        """infer.py:301-313: resize to target_size^2 (aspect not preserved), ToTensor, Normalize(0.5,0.5), network."""
        import cv2

        torch = self.torch
        if self.model is None:
            raise RuntimeError("EAST(model=...) is required for predict(): the detector network is outside this "
                               "package (use boxes_from_maps / predict_from_maps when the maps already exist)")
        resized = cv2.resize(img, (self.target_size, self.target_size))
        t = torch.from_numpy(np.ascontiguousarray(resized)).to(self.device).permute(2, 0, 1).float().div_(255.0)
        t = (t - 0.5) / 0.5
        with torch.no_grad():
            out = self.model(t.unsqueeze(0))
        return out["score"][0], out["geometry"][0]

    def predict_from_maps(self, score_map, geo_map, orig_hw, sort_reading_order=False):
        quads = self.boxes_from_maps(score_map, geo_map, orig_hw)
        words = []
        for quad in quads:
            pts = quad[:8].reshape(4, 2)
            words.append(Word(polygon=pts.tolist(), detection_confidence=float(quad[8])))  # infer.py:359-363
        if sort_reading_order and len(words) > 0:
            words = reorder_words(words)  # infer.py:366-385
        return Page(blocks=[Block(words=words)])

    def predict(self, img_or_path, vis=False, profile=False, return_maps=False, sort_reading_order=False):
        """Same contract as the reference's EAST.predict: dict with "page", "vis_image", "score_map", "geo_map"."""
        img = read_image(img_or_path)
        t0 = time.time()
        score, geo = self._forward(img)
        if profile:
            self.torch.cuda.synchronize()
            print(f"  Model inference: {time.time() - t0:.3f}s")
        t0 = time.time()
        page = self.predict_from_maps(score, geo, img.shape[:2], sort_reading_order=sort_reading_order)
        if profile:
            print(f"  Decode + NMS + box filters (device): {time.time() - t0:.3f}s")
            print(f"    Boxes after NMS: {sum(len(b.words) for b in page.blocks)}")
        vis_img = None
        if vis:
            raise NotImplementedError("visualisation is outside the hot path: pass the returned Page to the "
                                      "reference's visualize_page (the Page type is field-compatible)")
        return {
            "page": page,
            "vis_image": vis_img,
            "score_map": score[0].cpu().numpy() if return_maps else None,
            "geo_map": geo.cpu().numpy() if return_maps else None,
        }
