"""Generate tests/golden/letterbox.npz by EXECUTING the reference's utils.augmentations.letterbox (build container only).

    python -m oracle.make_golden_letterbox
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
CASES = [  # (h, w, kwargs) -- small images: the fixture stays a few hundred KB
    (60, 80, dict(new_shape=64)), (135, 101, dict(new_shape=96)), (47, 63, dict(new_shape=(52, 80), auto=False)),
    (38, 25, dict(new_shape=40, scaleup=False)), (42, 65, dict(new_shape=(48, 80), scaleFill=True, auto=False)),
    (64, 64, dict(new_shape=64, stride=64)), (63, 42, dict(new_shape=80, stride=16, color=(0, 127, 255))),
    (100, 37, dict(new_shape=(64, 64), auto=True, stride=32)), (31, 90, dict(new_shape=128)),
]


def main():
    sys.path.insert(0, str(ROOT))
    from oracle import refshim
    refshim.load()
    from utils.augmentations import letterbox          # the reference's function
    g = np.random.default_rng(5)
    arrs = {}
    for i, (h, w, kw) in enumerate(CASES):
        im = g.integers(0, 256, (h, w, 3), dtype=np.uint8)
        out, ratio, pad = letterbox(im, **kw)
        arrs[f'im{i}'] = im
        arrs[f'out{i}'] = out
        arrs[f'meta{i}'] = np.array([ratio[0], ratio[1], pad[0], pad[1]], dtype=np.float64)
    np.savez_compressed(ROOT / 'tests' / 'golden' / 'letterbox.npz', **arrs)
    print('letterbox.npz', sum(v.nbytes for v in arrs.values()) // 1024, 'KiB raw')


if __name__ == '__main__':
    main()
