"""Host-side image preparation of the detection path: `letterbox`, the mirror of
utils/augmentations.py:91-122 (same name, arguments and return triple).  It runs on the host like the reference's (the
image arrives from the decoder as a uint8 HWC array); the result goes to the GPU as uint8 and is normalised inside
`dmay_input_prep` (val.py:199-202)."""
from __future__ import annotations

import cv2
import numpy as np

__all__ = ['letterbox']


def letterbox(im, new_shape=(640, 640), color=(114, 114, 114), auto=True, scaleFill=False, scaleup=True, stride=32):
    """Resize `im` (HWC) to fit `new_shape` keeping the aspect ratio, then pad: to the smallest stride multiple (`auto`),
    not at all after a stretch (`scaleFill`), or to the full `new_shape`.  Returns (image, (w_ratio, h_ratio), (dw, dh))
    with dw, dh the padding PER SIDE (may be x.5: the extra pixel goes to the bottom / right)."""
    h0, w0 = im.shape[:2]
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    gain = min(new_shape[0] / h0, new_shape[1] / w0)
    if not scaleup:
        gain = min(gain, 1.0)
    ratio = (gain, gain)
    w1, h1 = int(round(w0 * gain)), int(round(h0 * gain))
    pad_w, pad_h = new_shape[1] - w1, new_shape[0] - h1
    if auto:
        pad_w, pad_h = np.mod(pad_w, stride), np.mod(pad_h, stride)
    elif scaleFill:
        pad_w, pad_h = 0.0, 0.0
        w1, h1 = new_shape[1], new_shape[0]
        ratio = (new_shape[1] / w0, new_shape[0] / h0)
    pad_w, pad_h = pad_w / 2, pad_h / 2
    if (w0, h0) != (w1, h1):
        im = cv2.resize(im, (w1, h1), interpolation=cv2.INTER_LINEAR)
    top, bottom = int(round(pad_h - 0.1)), int(round(pad_h + 0.1))
    left, right = int(round(pad_w - 0.1)), int(round(pad_w + 0.1))
    im = cv2.copyMakeBorder(im, top, bottom, left, right, cv2.BORDER_CONSTANT, value=color)
    return im, ratio, (pad_w, pad_h)
