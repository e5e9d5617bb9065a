"""TRBA recogniser wrapper: batch assembly on the B200, the reference's predict() contract.

Mirrors `TRBA._preprocess_image` + the batching loop of `TRBA.predict` (reference
recognizers/_trba/__init__.py:264-288, 382-390) and `get_val_transform` / `ResizeAndPadA`
(recognizers/_trba/data/transforms.py:62-120, 185-193).  The recogniser network (SEResNet31 + BiLSTM +
attention) is outside the hot path and is supplied by the caller: `model(batch (n,3,h,w) f32 on the device)`
must return a list of {"text", "confidence"} dicts (or (text, confidence) tuples) of length n.
"""
import numpy as np

from . import ops


class TRBA:
    def __init__(self, model=None, img_h=64, img_w=256, device=0, batch_size=32):
        # 64x256 is the reference class default (recognizers/_trba/__init__.py:150-151); configs/config.json
        # uses 32x128 -- both are supported
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("manuscript_b200.TRBA needs a CUDA device: there is no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", int(device)) if not isinstance(device, torch.device) else device
        self.model = model
        self.img_h, self.img_w = int(img_h), int(img_w)
        self.batch_size = int(batch_size)

    @staticmethod
    def _as_rgb(image):
        """recognizers/_trba/__init__.py:264-283 for array inputs (grey -> RGB, RGBA -> RGB)."""
        if not isinstance(image, np.ndarray):
            raise ValueError(f"Unsupported image type: {type(image)}")
        img = image
        if img.ndim == 2:
            img = np.repeat(img[:, :, None], 3, axis=2)
        elif img.shape[2] == 4:
            img = img[:, :, :3]
        if img.dtype != np.uint8:
            raise ValueError("TRBA expects uint8 images")
        return img

    def preprocess(self, images):
        """List of (h,w,3) uint8 crops -> (n,3,img_h,img_w) float32 CUDA tensor: what the reference builds with
        one cv2.resize + .to(device) per image followed by torch.stack."""
        torch = self.torch
        imgs = [self._as_rgb(im) for im in images]
        n = len(imgs)
        if n == 0:
            return torch.empty((0, 3, self.img_h, self.img_w), dtype=torch.float32, device=self.device)
        # pack the crops into one atlas so that a single kernel launch resamples all of them
        wmax = max(im.shape[1] for im in imgs)
        htot = sum(im.shape[0] for im in imgs)
        atlas = np.zeros((htot, wmax, 3), np.uint8)
        rects = np.zeros((n, 4), np.int32)
        y = 0
        for i, im in enumerate(imgs):
            h, w = im.shape[:2]
            atlas[y:y + h, :w] = im
            rects[i] = (0, y, w, y + h)
            y += h
        batch = ops.crop_resize_pad(atlas, rects, self.img_h, self.img_w)
        return torch.from_numpy(batch).to(self.device)

    def predict(self, images, batch_size=None, **model_kwargs):
        if not isinstance(images, list):
            images = [images]
        if self.model is None:
            raise RuntimeError("TRBA(model=...) is required for predict(): the recogniser network is outside this package")
        bs = int(batch_size or self.batch_size)
        results = []
        for i in range(0, len(images), bs):  # recognizers/_trba/__init__.py:382
            batch = self.preprocess(images[i:i + bs])
            results.extend(self.predict_batch(batch, **model_kwargs))
        return results

    def predict_batch(self, batch, **model_kwargs):
        """batch (n,3,h,w) f32 on the device -> list of {"text", "confidence"}."""
        if self.model is None:
            raise RuntimeError("TRBA(model=...) is required")
        out = self.model(batch, **model_kwargs)
        res = []
        for r in out:
            if isinstance(r, dict):
                res.append({"text": r.get("text", ""), "confidence": r.get("confidence")})
            elif isinstance(r, tuple) and len(r) == 2:
                res.append({"text": r[0], "confidence": r[1]})
            else:
                res.append({"text": str(r) if r is not None else "", "confidence": None})
        return res
