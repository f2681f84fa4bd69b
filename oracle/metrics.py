"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — numpy restatement of the reference's mAP arithmetic:
`process_batch` (val.py:62-83), `box_iou` (utils/metrics.py:254-276), `compute_ap` (utils/metrics.py:86-111),
`ap_per_class` (utils/metrics.py:21-83) and the statistics tail of `val.run` (val.py:236-288).
Pinned by tests/golden/metrics.npz (outputs of the executed reference, oracle/make_golden.py)."""
from __future__ import annotations

import numpy as np


def box_iou(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    wh = np.clip(np.minimum(a[:, None, 2:], b[None, :, 2:]) - np.maximum(a[:, None, :2], b[None, :, :2]), 0, None)
    inter = wh[..., 0] * wh[..., 1]
    with np.errstate(divide='ignore', invalid='ignore'):      # degenerate boxes: 0/0 -> nan, like the reference
        return inter / (area_a[:, None] + area_b[None, :] - inter)


def process_batch(det, labels, iouv):
    """det [N,6] xyxy,conf,cls ; labels [M,5] cls,xyxy -> bool [N, len(iouv)]."""
    iouv = np.asarray(iouv, np.float32)
    correct = np.zeros((det.shape[0], iouv.shape[0]), bool)
    iou = box_iou(labels[:, 1:], det[:, :4])
    li, di = np.nonzero((iou >= iouv[0]) & (labels[:, 0:1] == det[None, :, 5]))
    if li.size:
        m = np.stack([li.astype(np.float64), di.astype(np.float64), iou[li, di].astype(np.float32).astype(np.float64)], 1)
        if li.size > 1:
            m = m[m[:, 2].argsort()[::-1]]
            m = m[np.unique(m[:, 1], return_index=True)[1]]
            m = m[np.unique(m[:, 0], return_index=True)[1]]
        correct[m[:, 1].astype(int)] = m[:, 2:3].astype(np.float32) >= iouv
    return correct


def compute_ap(recall, precision):
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.concatenate(([1.0], precision, [0.0]))
    for i in range(mpre.size - 2, -1, -1):          # precision envelope, explicit loop
        mpre[i] = max(mpre[i], mpre[i + 1])
    x = np.linspace(0, 1, 101)
    y = np.interp(x, mrec, mpre)
    return float(((y[1:] + y[:-1]) / 2 * (x[1:] - x[:-1])).sum())


def ap_per_class(tp, conf, pred_cls, target_cls):
    order = np.argsort(-conf)
    tp, conf, pred_cls = tp[order], conf[order], pred_cls[order]
    classes = np.unique(target_cls)
    px = np.linspace(0, 1, 1000)
    ap = np.zeros((classes.size, tp.shape[1]))
    p = np.zeros((classes.size, 1000))
    r = np.zeros((classes.size, 1000))
    for ci, c in enumerate(classes):
        sel = pred_cls == c
        n_l = int((target_cls == c).sum())
        if not sel.any() or n_l == 0:
            continue
        tpc = tp[sel].cumsum(0)
        fpc = (1 - tp[sel]).cumsum(0)
        recall = tpc / (n_l + 1e-16)
        precision = tpc / (tpc + fpc)
        r[ci] = np.interp(-px, -conf[sel], recall[:, 0], left=0)
        p[ci] = np.interp(-px, -conf[sel], precision[:, 0], left=1)
        for j in range(tp.shape[1]):
            ap[ci, j] = compute_ap(recall[:, j], precision[:, j])
    f1 = 2 * p * r / (p + r + 1e-16)
    i = f1.mean(0).argmax()
    return p[:, i], r[:, i], ap, f1[:, i], classes.astype('int32')


def evaluate(dets, labels, img_hw):
    """dets: list of [n,6] xyxy,conf,cls (pixels) per image; labels: list of [m,5] cls + xywh NORMALISED per image.
    -> (mp, mr, map50, map) exactly as val.run accumulates them (no letterbox: ratio 1, pad 0)."""
    h, w = img_hw
    iouv = np.linspace(0.5, 0.95, 10).astype(np.float32)
    stats = []
    for d, l in zip(dets, labels):
        d = np.asarray(d, np.float32)
        l = np.asarray(l, np.float32).reshape(-1, 5)
        tcls = l[:, 0].tolist()
        if d.shape[0] == 0:
            if l.shape[0]:
                stats.append((np.zeros((0, 10), bool), np.zeros(0, np.float32), np.zeros(0, np.float32), tcls))
            continue
        if l.shape[0]:
            xywh = l[:, 1:5] * np.array([w, h, w, h], np.float32)
            half = xywh[:, 2:] / 2
            lab = np.concatenate([l[:, 0:1], xywh[:, :2] - half, xywh[:, :2] + half], 1)
            dd = d.copy()
            dd[:, [0, 2]] = dd[:, [0, 2]].clip(0, w)
            dd[:, [1, 3]] = dd[:, [1, 3]].clip(0, h)
            lab[:, [1, 3]] = lab[:, [1, 3]].clip(0, w)
            lab[:, [2, 4]] = lab[:, [2, 4]].clip(0, h)
            correct = process_batch(dd, lab, iouv)
        else:
            correct = np.zeros((d.shape[0], 10), bool)
        stats.append((correct, d[:, 4], d[:, 5], tcls))
    if not stats:
        return 0.0, 0.0, 0.0, 0.0
    tp, conf, pcls, tcls = [np.concatenate([np.asarray(t) for t in x], 0) for x in zip(*stats)]
    if not tp.any():
        return 0.0, 0.0, 0.0, 0.0
    p, r, ap, f1, _ = ap_per_class(tp, conf, pcls, tcls)
    return float(p.mean()), float(r.mean()), float(ap[:, 0].mean()), float(ap.mean(1).mean())
