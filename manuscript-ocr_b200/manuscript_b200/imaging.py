"""Image input and page visualisation at the boundary of the path (host-side helpers, no kernels).

`read_image` has the contract of the reference's detectors/_east/utils.py:477-497 (path -> RGB uint8 array through
OpenCV with a PIL second try, arrays pass through, anything else is a TypeError) and additionally accepts PIL images,
which the reference's Pipeline.predict signature advertises (_pipeline.py:58).

`visualize_page` stands in for the reference's visualize_page (utils.py:95-230) where Pipeline.predict(vis=True) and
EAST.predict(vis=True) return a PIL image: polygon outlines, and with show_order=True the reading-order numbers joined
by a line.  It draws with PIL only and does not try to reproduce the reference's pixels (visualisation is not on the
measured path); the Page type is field-compatible, so the reference's own visualize_page can be used on the result.
"""
from pathlib import Path

import numpy as np


def _is_pil(obj):
    try:
        from PIL import Image
    except Exception:  # pragma: no cover - PIL is part of the image
        return False
    return isinstance(obj, Image.Image)


def read_image(source):
    """str / Path -> (H, W, 3) uint8 RGB array; ndarray -> itself; PIL image -> RGB array.
    FileNotFoundError when neither OpenCV nor PIL can decode the file, TypeError for other inputs."""
    if isinstance(source, np.ndarray):
        return source
    if _is_pil(source):
        return np.asarray(source.convert("RGB"))
    if not isinstance(source, (str, Path)):
        raise TypeError(f"Unsupported type for image input: {type(source)}")
    import cv2

    path = str(source)
    bgr = cv2.imread(path)
    if bgr is not None:
        return cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
    try:
        from PIL import Image

        with Image.open(path) as im:
            return np.array(im.convert("RGB"))
    except Exception as exc:
        raise FileNotFoundError(f"Cannot read image with cv2 or PIL: {source}. Error: {exc}")


def visualize_page(image, page, *, show_order=False, color=(0, 0, 255), thickness=2, line_color=(0, 255, 0),
                   number_color=(255, 255, 255), number_bg=(0, 0, 0)):
    """PIL image with every word polygon of `page` outlined; show_order=True numbers the words in list order (the
    reading order once Pipeline.predict has sorted them) and joins their centres."""
    from PIL import Image, ImageDraw

    if _is_pil(image):
        canvas = image.convert("RGB").copy()
    else:
        arr = np.asarray(image)
        if arr.ndim == 2:
            arr = np.repeat(arr[:, :, None], 3, axis=2)
        canvas = Image.fromarray(np.ascontiguousarray(arr[:, :, :3]).astype(np.uint8))
    draw = ImageDraw.Draw(canvas)
    centres = []
    for block in page.blocks:
        for word in block.words:
            pts = [(float(x), float(y)) for x, y in word.polygon]
            if len(pts) >= 2:
                draw.line(pts + [pts[0]], fill=tuple(color), width=int(thickness))
                centres.append((sum(p[0] for p in pts) / len(pts), sum(p[1] for p in pts) / len(pts)))
    if show_order and centres:
        if len(centres) > 1:
            draw.line(centres, fill=tuple(line_color), width=max(1, int(thickness) // 2))
        for k, (cx, cy) in enumerate(centres, 1):
            label = str(k)
            box = draw.textbbox((cx, cy), label)
            draw.rectangle((box[0] - 1, box[1] - 1, box[2] + 1, box[3] + 1), fill=tuple(number_bg))
            draw.text((cx, cy), label, fill=tuple(number_color))
    return canvas
