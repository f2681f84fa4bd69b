"""Generate tests/golden/metrics.npz by EXECUTING the reference's mAP code (build container only):
`val.process_batch` (val.py:62-83) and `utils.metrics.ap_per_class` (utils/metrics.py:21-83) on a seeded synthetic
set of detections and labels (6 images, 5 classes, ties and unmatched classes included).

    python -m oracle.make_golden_metrics
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent


def synth(seed=3, n_img=6, nc=5, S=160):
    g = np.random.default_rng(seed)
    dets, labels = [], []
    for i in range(n_img):
        m = int(g.integers(0, 7)) if i != 2 else 0           # image 2 has no labels
        cxy = g.uniform(20, S - 20, (m, 2))
        wh = g.uniform(8, 40, (m, 2))
        cls = g.integers(0, nc - 1, (m, 1))                    # class nc-1 never labelled
        labels.append(np.concatenate([cls, cxy / S, wh / S], 1).astype(np.float32))
        n = int(g.integers(0, 30)) if i != 4 else 0            # image 4 has no detections
        d = []
        for k in range(n):
            if m and g.random() < 0.6:                         # near a label
                j = int(g.integers(0, m))
                c = cxy[j] + g.normal(0, 2.5, 2)
                s = wh[j] * g.uniform(0.8, 1.25, 2)
                cl = cls[j, 0] if g.random() < 0.8 else g.integers(0, nc)
            else:
                c, s, cl = g.uniform(10, S - 10, 2), g.uniform(6, 50, 2), g.integers(0, nc)
            conf = np.round(g.uniform(0.01, 1.0), 2)           # two-decimal scores: ties across images
            d.append([c[0] - s[0] / 2, c[1] - s[1] / 2, c[0] + s[0] / 2, c[1] + s[1] / 2, conf, cl])
        d = np.array(d, np.float32).reshape(-1, 6)
        dets.append(d[np.argsort(-d[:, 4], kind='stable')] if n else d)
    return dets, labels, S, nc


def main():
    sys.path.insert(0, str(ROOT))
    from oracle import refshim
    refshim.load()
    import val as V                      # the reference's val.py
    from utils.general import xywh2xyxy
    from utils.metrics import ap_per_class
    dets, labels, S, nc = synth()
    iouv = torch.linspace(0.5, 0.95, 10)
    stats, arrs = [], {}
    for i, (d, l) in enumerate(zip(dets, labels)):
        arrs[f'det_{i}'], arrs[f'lab_{i}'] = d, l
        dt, lt = torch.from_numpy(d), torch.from_numpy(l)
        tcls = lt[:, 0].tolist() if len(lt) else []
        if len(dt) == 0:
            if len(lt):
                stats.append((torch.zeros(0, 10, dtype=torch.bool), torch.Tensor(), torch.Tensor(), tcls))
            continue
        if len(lt):
            tbox = xywh2xyxy(lt[:, 1:5] * S)
            correct = V.process_batch(dt, torch.cat((lt[:, 0:1], tbox), 1), iouv)
        else:
            correct = torch.zeros(dt.shape[0], 10, dtype=torch.bool)
        arrs[f'correct_{i}'] = correct.numpy()
        stats.append((correct, dt[:, 4], dt[:, 5], tcls))
    st = [np.concatenate(x, 0) for x in zip(*stats)]
    p, r, ap, f1, ap_class = ap_per_class(*st, plot=False, names={})   # names must be a dict (utils/metrics.py:74)
    arrs.update(p=p, r=r, ap=ap, f1=f1, ap_class=ap_class, S=np.array(S), nc=np.array(nc), n_img=np.array(len(dets)),
                summary=np.array([p.mean(), r.mean(), ap[:, 0].mean(), ap.mean(1).mean()]))
    np.savez_compressed(ROOT / 'tests' / 'golden' / 'metrics.npz', **arrs)
    print('metrics.npz', arrs['summary'])


if __name__ == '__main__':
    main()
