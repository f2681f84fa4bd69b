"""Seeded synthetic labelled detection set (test infrastructure; numpy only so that the GPU box regenerates the
exact bytes the build container trained / evaluated on — PCG64 integer and double streams are platform independent).

Images: S x S uint8, a smooth low-contrast background plus 2..9 filled shapes; the class is (shape, colour family):
0 = red rectangle, 1 = green rectangle, 2 = blue ellipse, 3 = yellow ellipse.  Labels: (cls, xc, yc, w, h) / S as in
the reference's dataloader targets (val.py:199-206: targets[n,6] = image index + label row).
"""
from __future__ import annotations

import numpy as np

NC = 4
_BASE = np.array([[200, 40, 40], [40, 180, 60], [50, 70, 210], [220, 200, 40]], np.int64)


def make_image(g: np.random.Generator, S: int = 640, max_obj: int = 9):
    """-> (img uint8 [3,S,S], labels float32 [n,5])."""
    # background: 16x16 blocks of gray, bilinear-free (nearest) upsampling + per-pixel integer noise
    nb = S // 16 + 1
    coarse = g.integers(90, 150, (nb, nb))
    bg = np.repeat(np.repeat(coarse, 16, 0), 16, 1)[:S, :S]
    img = np.stack([bg, bg, bg], 0).astype(np.int64)
    img += g.integers(-12, 13, (3, S, S))
    n = int(g.integers(2, max_obj + 1))
    labels, boxes = [], []
    yy, xx = np.mgrid[0:S, 0:S]
    tries = 0
    while len(labels) < n and tries < 50:
        tries += 1
        w = int(g.integers(S // 20, S // 4))
        h = int(g.integers(S // 20, S // 4))
        x0 = int(g.integers(2, S - w - 2))
        y0 = int(g.integers(2, S - h - 2))
        # keep overlaps small: a new object may cover at most 20 % of an earlier one and vice versa
        ok = True
        for (a0, b0, a1, b1) in boxes:
            iw = min(x0 + w, a1) - max(x0, a0)
            ih = min(y0 + h, b1) - max(y0, b0)
            if iw > 0 and ih > 0 and iw * ih > 0.2 * min(w * h, (a1 - a0) * (b1 - b0)):
                ok = False
                break
        if not ok:
            continue
        c = int(g.integers(0, NC))
        col = _BASE[c] + g.integers(-25, 26, 3)
        if c < 2:
            mask = (xx >= x0) & (xx < x0 + w) & (yy >= y0) & (yy < y0 + h)
        else:
            cx, cy = x0 + w / 2.0, y0 + h / 2.0
            mask = ((xx + 0.5 - cx) / (w / 2.0)) ** 2 + ((yy + 0.5 - cy) / (h / 2.0)) ** 2 <= 1.0
        for k in range(3):
            img[k][mask] = col[k] + g.integers(-8, 9, int(mask.sum()))
        boxes.append((x0, y0, x0 + w, y0 + h))
        labels.append([c, (x0 + w / 2.0) / S, (y0 + h / 2.0) / S, w / S, h / S])
    return np.clip(img, 0, 255).astype(np.uint8), np.asarray(labels, np.float32).reshape(-1, 5)


def make_batch(seed: int, B: int, S: int = 640):
    """-> (imgs uint8 [B,3,S,S], targets float32 [n,6] = (image index, cls, xc, yc, w, h))."""
    g = np.random.default_rng(seed)
    imgs, tg = [], []
    for i in range(B):
        im, lab = make_image(g, S)
        imgs.append(im)
        tg.append(np.concatenate([np.full((len(lab), 1), i, np.float32), lab], 1))
    return np.stack(imgs, 0), np.concatenate(tg, 0).astype(np.float32)
