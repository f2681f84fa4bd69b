"""Post-processing of the detection path with the reference's call signatures (utils/general.py).

`non_max_suppression` keeps the reference signature and return type; on CUDA tensors the whole batch
runs on our kernels (filter -> sort -> greedy NMS, see ops.nms_batched); on CPU tensors the reference's
torch/torchvision body runs (there are no CPU kernels in this package).
"""
from __future__ import annotations

import logging
import math
import time

import numpy as np
import torch

LOGGER = logging.getLogger("dma_yolo_b200")


def make_divisible(x, divisor):
    """utils/general.py:450-452"""
    return math.ceil(x / divisor) * divisor


def check_version(current='0.0.0', minimum='0.0.0', name='version ', pinned=False, hard=False):
    """utils/general.py:244-251 without pkg_resources."""
    def parse(v):
        out = []
        for tok in str(v).split('+')[0].split('.'):
            num = ''.join(ch for ch in tok if ch.isdigit())
            out.append(int(num) if num else 0)
        return tuple(out)
    c, m = parse(current), parse(minimum)
    result = (c == m) if pinned else (c >= m)
    if hard:
        assert result, f'{name}{minimum} required, but {name}{current} is currently installed'
    return result


def xyxy2xywh(x):
    y = x.clone() if isinstance(x, torch.Tensor) else np.copy(x)
    y[:, 0] = (x[:, 0] + x[:, 2]) / 2
    y[:, 1] = (x[:, 1] + x[:, 3]) / 2
    y[:, 2] = x[:, 2] - x[:, 0]
    y[:, 3] = x[:, 3] - x[:, 1]
    return y


def xywh2xyxy(x):
    """utils/general.py:539-546"""
    y = x.clone() if isinstance(x, torch.Tensor) else np.copy(x)
    y[:, 0] = x[:, 0] - x[:, 2] / 2
    y[:, 1] = x[:, 1] - x[:, 3] / 2
    y[:, 2] = x[:, 0] + x[:, 2] / 2
    y[:, 3] = x[:, 1] + x[:, 3] / 2
    return y


def clip_coords(boxes, shape):
    """utils/general.py:621-630"""
    if isinstance(boxes, torch.Tensor):
        boxes[:, 0].clamp_(0, shape[1])
        boxes[:, 1].clamp_(0, shape[0])
        boxes[:, 2].clamp_(0, shape[1])
        boxes[:, 3].clamp_(0, shape[0])
    else:
        boxes[:, [0, 2]] = boxes[:, [0, 2]].clip(0, shape[1])
        boxes[:, [1, 3]] = boxes[:, [1, 3]].clip(0, shape[0])


def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None):
    """utils/general.py:605-618"""
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    coords[:, [0, 2]] -= pad[0]
    coords[:, [1, 3]] -= pad[1]
    coords[:, :4] /= gain
    clip_coords(coords, img0_shape)
    return coords


def _nms_torch_cpu(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, labels, max_det):
    """Reference body (utils/general.py:645-725) for non-CUDA inputs; same third-party op
    (torchvision.ops.nms) as the reference.  The wall-clock `time_limit` break is not reproduced."""
    import torchvision
    nc = prediction.shape[2] - 5
    xc = prediction[..., 4] > conf_thres
    max_wh, max_nms = 4096, 30000
    multi_label &= nc > 1
    output = [torch.zeros((0, 6), device=prediction.device)] * prediction.shape[0]
    for xi, x in enumerate(prediction):
        x = x[xc[xi]]
        if labels and len(labels[xi]):
            l = labels[xi]
            v = torch.zeros((len(l), nc + 5), device=x.device)
            v[:, :4] = l[:, 1:5]
            v[:, 4] = 1.0
            v[range(len(l)), l[:, 0].long() + 5] = 1.0
            x = torch.cat((x, v), 0)
        if not x.shape[0]:
            continue
        x[:, 5:] *= x[:, 4:5]
        box = xywh2xyxy(x[:, :4])
        if multi_label:
            i, j = (x[:, 5:] > conf_thres).nonzero(as_tuple=False).T
            x = torch.cat((box[i], x[i, j + 5, None], j[:, None].float()), 1)
        else:
            conf, j = x[:, 5:].max(1, keepdim=True)
            x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]
        if classes is not None:
            x = x[(x[:, 5:6] == torch.tensor(classes, device=x.device)).any(1)]
        n = x.shape[0]
        if not n:
            continue
        elif n > max_nms:
            x = x[x[:, 4].argsort(descending=True, stable=True)[:max_nms]]
        c = x[:, 5:6] * (0 if agnostic else max_wh)
        boxes, scores = x[:, :4] + c, x[:, 4]
        i = torchvision.ops.nms(boxes, scores, iou_thres)
        if i.shape[0] > max_det:
            i = i[:max_det]
        output[xi] = x[i]
    return output


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        labels=(), max_det=300):
    """Runs Non-Maximum Suppression (NMS) on inference results — utils/general.py:633-725.

    Returns: list of detections, one (n,6) tensor per image [xyxy, conf, cls], descending confidence.
    """
    assert 0 <= conf_thres <= 1, f'Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0'
    assert 0 <= iou_thres <= 1, f'Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0'
    if not prediction.is_cuda:
        return _nms_torch_cpu(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, labels, max_det)

    from .. import ops
    from ..lazy import LazyPred
    kw = dict(classes=classes, agnostic=agnostic, multi_label=multi_label, max_det=max_det)
    if labels and any(len(l) for l in labels):
        # val.py --save-hybrid: a-priori label rows are appended per image, so row counts differ ->
        # one launch sequence per image over [pred_i ; label rows] (same kernels).
        dense = prediction.dense() if isinstance(prediction, LazyPred) else prediction
        nc = dense.shape[2] - 5
        res = []
        for xi in range(dense.shape[0]):
            x = dense[xi].float()
            l = labels[xi] if xi < len(labels) else ()
            if len(l):
                v = torch.zeros((len(l), nc + 5), device=x.device)
                v[:, :4] = l[:, 1:5]
                v[:, 4] = 1.0
                v[range(len(l)), l[:, 0].long() + 5] = 1.0
                x = torch.cat((x, v), 0)
            out, cnt = ops.nms_batched(x[None].contiguous(), conf_thres, iou_thres, **kw)
            res.append(out[0, :int(cnt[0].item())])
        return res
    if isinstance(prediction, LazyPred) and prediction._dense is None:
        out, cnt, packed = ops.nms_batched(None, conf_thres, iou_thres, levels=prediction._levels, na=prediction._na,
                                           nc=prediction._no - 5, return_packed=True, **kw)
    else:
        dense = prediction.dense() if isinstance(prediction, LazyPred) else prediction
        out, cnt, packed = ops.nms_batched(dense, conf_thres, iou_thres, return_packed=True, **kw)
    return Detections(out, cnt, packed)


class Detections(list):
    """The reference's return type — a list of (n,6) tensors, one per image — plus the batch buffers its elements are
    views of: `padded` [N, max_det, 6] fp32, `counts` [N] int32 and `packed` (the flat buffer both live in) on the
    device, so one D2H / one all_gather moves the batch.  The list elements need the per-image counts on the host:
    that one D2H of N ints happens on first access to an element (iteration, indexing, comparison ...), not inside
    `non_max_suppression` — a caller that only forwards `padded` / `counts` / `packed` never synchronises."""

    def __init__(self, padded, counts, packed=None):
        super().__init__([None] * padded.shape[0])
        self.padded, self.counts, self.packed = padded, counts, packed
        self._pending = True

    def _fill(self):
        if self._pending:
            self._pending = False
            c = self.counts.tolist()
            for i, n in enumerate(c):
                list.__setitem__(self, i, self.padded[i, :n])
        return self

    def __iter__(self):
        return list.__iter__(self._fill())

    def __getitem__(self, i):
        return list.__getitem__(self._fill(), i)

    def __reversed__(self):
        return list.__reversed__(self._fill())

    def __contains__(self, v):
        return list.__contains__(self._fill(), v)

    def __eq__(self, o):
        return list.__eq__(self._fill(), o)

    __hash__ = None

    def __add__(self, o):
        return list(self._fill()) + list(o)

    def __repr__(self):
        return list.__repr__(self._fill())

    def copy(self):
        return list(self._fill())

    def __reduce__(self):
        return (list, (list(self._fill()),))
