"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — numpy restatement of the reference NMS path.

`nms_reference` restates torchvision.ops.nms (CPU kernel `nms_kernel_impl<float>`, the third-party
arithmetic behind utils/general.py:708; torchvision is not under /root/reference — requirements.txt:12
`torchvision>=0.8.1`, installed 0.26.0): stable descending sort, fp32 areas / intersection / IoU with
true division, `iou > threshold` evaluated in double.  `non_max_suppression` restates
utils/general.py:633-725 around it.  All arithmetic is np.float32 scalar/array arithmetic so that the
rounding of every intermediate matches the reference.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def nms_reference(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float, max_keep: int | None = None) -> np.ndarray:
    """Greedy NMS; returns kept indices in descending-score order (ties: ascending index)."""
    boxes = np.ascontiguousarray(boxes, dtype=f32)
    scores = np.ascontiguousarray(scores, dtype=f32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)                      # fp32
    order = np.argsort(-scores.astype(np.float64), kind='stable')  # stable descending
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr = float(iou_threshold)                         # python double
    with np.errstate(divide='ignore', invalid='ignore'):
        for _i in range(n):
            i = order[_i]
            if suppressed[i]:
                continue
            keep.append(i)
            if max_keep is not None and len(keep) >= max_keep:
                break
            rest = order[_i + 1:]
            xx1 = np.maximum(x1[i], x1[rest])
            yy1 = np.maximum(y1[i], y1[rest])
            xx2 = np.minimum(x2[i], x2[rest])
            yy2 = np.minimum(y2[i], y2[rest])
            w = np.maximum(f32(0), xx2 - xx1)
            h = np.maximum(f32(0), yy2 - yy1)
            inter = w * h
            ovr = inter / (areas[i] + areas[rest] - inter)        # fp32 true division (0/0 -> nan -> kept)
            suppressed[rest[ovr.astype(np.float64) > thr]] = True
    return np.asarray(keep, dtype=np.int64)


def xywh2xyxy(x: np.ndarray) -> np.ndarray:
    """utils/general.py:539-546"""
    y = x.copy()
    y[:, 0] = x[:, 0] - x[:, 2] / f32(2)
    y[:, 1] = x[:, 1] - x[:, 3] / f32(2)
    y[:, 2] = x[:, 0] + x[:, 2] / f32(2)
    y[:, 3] = x[:, 1] + x[:, 3] / f32(2)
    return y


def non_max_suppression(prediction: np.ndarray, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        multi_label=False, labels=(), max_det=300, max_nms=30000):
    """utils/general.py:633-725 on a [B, R, 5+nc] fp32 array -> list of (n,6) fp32 arrays.
    Deviations (documented): no wall-clock time_limit; the >max_nms truncation uses a
    STABLE descending sort (the reference's argsort is unstable, i.e. implementation-defined on ties)."""
    prediction = np.asarray(prediction, dtype=f32)
    nc = prediction.shape[2] - 5
    max_wh = 4096          # max_nms: 30000 in the reference (utils/general.py:656); a parameter here so that tests can
                           # exercise the truncation path on small inputs
    multi_label = bool(multi_label) and nc > 1
    thr = f32(conf_thres) if False else conf_thres  # comparisons: fp32 tensor vs python scalar -> fp32 compare
    out = []
    for xi in range(prediction.shape[0]):
        x = prediction[xi]
        x = x[x[:, 4] > f32(conf_thres)].copy()
        if labels and len(labels[xi]):
            l = np.asarray(labels[xi], dtype=f32)
            v = np.zeros((len(l), nc + 5), dtype=f32)
            v[:, :4] = l[:, 1:5]
            v[:, 4] = 1.0
            v[np.arange(len(l)), l[:, 0].astype(np.int64) + 5] = 1.0
            x = np.concatenate((x, v), 0)
        if not x.shape[0]:
            out.append(np.zeros((0, 6), dtype=f32))
            continue
        x[:, 5:] *= x[:, 4:5]
        box = xywh2xyxy(x[:, :4])
        if multi_label:
            i, j = np.nonzero(x[:, 5:] > f32(conf_thres))
            x = np.concatenate((box[i], x[i, j + 5, None], j[:, None].astype(f32)), 1)
        else:
            j = np.argmax(x[:, 5:], axis=1)          # first maximum
            conf = x[np.arange(len(x)), j + 5]
            x = np.concatenate((box, conf[:, None], j[:, None].astype(f32)), 1)[conf > f32(conf_thres)]
        if classes is not None:
            x = x[np.isin(x[:, 5], np.asarray(classes, dtype=f32))]
        n = x.shape[0]
        if not n:
            out.append(np.zeros((0, 6), dtype=f32))
            continue
        elif n > max_nms:
            x = x[np.argsort(-x[:, 4].astype(np.float64), kind='stable')[:max_nms]]
        c = x[:, 5:6] * f32(0 if agnostic else max_wh)
        boxes, scores = x[:, :4] + c, x[:, 4]
        i = nms_reference(boxes, scores, iou_thres, max_keep=max_det)
        out.append(x[i[:max_det]])
    return out
