"""Host-side detection metrics with the reference's names and signatures (utils/metrics.py).

These stay numpy / torch on the host exactly like the reference (SURVEY.md 8f-3 lists a device version as a later
row): the B200 path ends at the per-image detections, which `val.run` feeds into `process_batch` and
`ap_per_class` unchanged.  Plotting arguments are accepted and ignored (matplotlib is not a dependency)."""
from __future__ import annotations

import numpy as np
import torch


def box_iou(box1, box2):
    """utils/metrics.py:254-276 — pairwise IoU of (x1, y1, x2, y2) boxes, [N, M]."""
    area1 = (box1[:, 2] - box1[:, 0]) * (box1[:, 3] - box1[:, 1])
    area2 = (box2[:, 2] - box2[:, 0]) * (box2[:, 3] - box2[:, 1])
    lt = torch.max(box1[:, None, :2], box2[:, :2])
    rb = torch.min(box1[:, None, 2:], box2[:, 2:])
    inter = (rb - lt).clamp(0).prod(2)
    return inter / (area1[:, None] + area2 - inter)


def compute_ap(recall, precision):
    """utils/metrics.py:86-111 — 101-point interpolated AP of the precision envelope."""
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.concatenate(([1.0], precision, [0.0]))
    mpre = np.flip(np.maximum.accumulate(np.flip(mpre)))
    x = np.linspace(0, 1, 101)
    y = np.interp(x, mrec, mpre)
    ap = float(np.sum((y[1:] + y[:-1]) * 0.5 * np.diff(x)))   # == np.trapz(y, x)
    return ap, mpre, mrec


def ap_per_class(tp, conf, pred_cls, target_cls, plot=False, save_dir='.', names=()):
    """utils/metrics.py:21-83 — per-class P, R, AP[nc, niou], F1 at the max-mean-F1 confidence, classes."""
    order = np.argsort(-conf)
    tp, conf, pred_cls = tp[order], conf[order], pred_cls[order]
    unique_classes = np.unique(target_cls)
    nc = unique_classes.shape[0]
    px = np.linspace(0, 1, 1000)
    ap, p, r = np.zeros((nc, tp.shape[1])), np.zeros((nc, 1000)), np.zeros((nc, 1000))
    for ci, c in enumerate(unique_classes):
        sel = pred_cls == c
        n_l = (target_cls == c).sum()
        if sel.sum() == 0 or n_l == 0:
            continue
        fpc = (1 - tp[sel]).cumsum(0)
        tpc = tp[sel].cumsum(0)
        recall = tpc / (n_l + 1e-16)
        r[ci] = np.interp(-px, -conf[sel], recall[:, 0], left=0)
        precision = tpc / (tpc + fpc)
        p[ci] = np.interp(-px, -conf[sel], precision[:, 0], left=1)
        for j in range(tp.shape[1]):
            ap[ci, j] = compute_ap(recall[:, j], precision[:, j])[0]
    f1 = 2 * p * r / (p + r + 1e-16)
    i = f1.mean(0).argmax()
    return p[:, i], r[:, i], ap, f1[:, i], unique_classes.astype('int32')
