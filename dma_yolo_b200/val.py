"""val.py-style evaluation on the B200 path: forward -> fused decode/filter -> batched NMS on the GPU, the
detections of all ranks all-gathered (NCCL), then the reference's host-side statistics unchanged.

`process_batch` and the statistics block follow val.py:62-83 and val.py:236-288; `run` accepts the in-memory
dataloader form that the reference's `val.run(model=..., dataloader=...)` call takes (the file-based data pipeline is
out of scope)."""
from __future__ import annotations

import numpy as np
import torch

from .utils.general import non_max_suppression, scale_coords, xywh2xyxy
from .utils.metrics import ap_per_class, box_iou


def process_batch(detections, labels, iouv):
    """val.py:62-83 — correct[N, niou]: detection i matches a label of its class at IoU level j.
    detections [N,6] = x1,y1,x2,y2,conf,cls ; labels [M,5] = cls,x1,y1,x2,y2."""
    correct = torch.zeros(detections.shape[0], iouv.shape[0], dtype=torch.bool, device=iouv.device)
    iou = box_iou(labels[:, 1:], detections[:, :4])
    x = torch.where((iou >= iouv[0]) & (labels[:, 0:1] == detections[:, 5]))
    if x[0].shape[0]:
        matches = torch.cat((torch.stack(x, 1), iou[x[0], x[1]][:, None]), 1).cpu().numpy()   # [label, detection, iou]
        if x[0].shape[0] > 1:
            matches = matches[matches[:, 2].argsort()[::-1]]
            matches = matches[np.unique(matches[:, 1], return_index=True)[1]]
            matches = matches[np.unique(matches[:, 0], return_index=True)[1]]
        matches = torch.Tensor(matches).to(iouv.device)
        correct[matches[:, 1].long()] = matches[:, 2:3] >= iouv
    return correct


@torch.no_grad()
def run(data, model=None, dataloader=None, batch_size=32, imgsz=640, conf_thres=0.001, iou_thres=0.6, single_cls=False,
        half=False, plots=False, rank=0, world_size=1, **_ignored):
    """-> ((mp, mr, map50, map), maps[nc]) like the head of the reference's return value (val.py:333-347).
    `dataloader` yields (img uint8/float [B,3,H,W], targets [n,6] = (img_idx, cls, xywh normalised), paths, shapes);
    with world_size > 1 every rank passes ITS slice of each batch and rank 0 returns the statistics."""
    from .dist import all_gather_detections, pad_detections, unpad
    assert model is not None and dataloader is not None, 'pass model= and an in-memory dataloader='
    device = next(model.parameters()).device
    model.eval()
    nc = 1 if single_cls else int(data['nc'])
    iouv = torch.linspace(0.5, 0.95, 10)
    stats = []
    max_det = 300
    for img, targets, paths, shapes in dataloader:
        img = img.to(device, non_blocking=True)
        if img.dtype != torch.uint8 or device.type != 'cuda':
            img = img.float() / 255          # val.py:199-202; uint8 on CUDA is normalised inside the prep kernel
        nb, _, height, width = img.shape
        out, _ = model(img)
        targets = targets.clone()
        targets[:, 2:] *= torch.tensor([width, height, width, height], dtype=targets.dtype)
        out = non_max_suppression(out, conf_thres, iou_thres, multi_label=True, agnostic=single_cls, max_det=max_det)
        if world_size > 1:   # rank-then-image order (np.argsort in ap_per_class is order sensitive)
            if getattr(out, 'packed', None) is not None:     # the NMS output buffer is already the exchange layout
                out = unpad(*all_gather_detections(out.padded, out.counts, packed=out.packed))
            else:
                p, c = pad_detections(out, max_det, device)
                out = unpad(*all_gather_detections(p, c))
            if rank != 0:
                continue
        for si, pred in enumerate(out):
            pred = pred.float().cpu()
            labels = targets[targets[:, 0] == si, 1:]
            nl = len(labels)
            tcls = labels[:, 0].tolist() if nl else []
            if len(pred) == 0:
                if nl:
                    stats.append((torch.zeros(0, iouv.numel(), dtype=torch.bool), torch.Tensor(), torch.Tensor(), tcls))
                continue
            if single_cls:
                pred[:, 5] = 0
            predn = pred.clone()
            shape = shapes[si][0] if shapes is not None else (height, width)
            ratio_pad = shapes[si][1] if shapes is not None else None
            scale_coords((height, width), predn[:, :4], shape, ratio_pad)
            if nl:
                tbox = xywh2xyxy(labels[:, 1:5])
                scale_coords((height, width), tbox, shape, ratio_pad)
                correct = process_batch(predn, torch.cat((labels[:, 0:1], tbox), 1), iouv)
            else:
                correct = torch.zeros(pred.shape[0], iouv.numel(), dtype=torch.bool)
            stats.append((correct, pred[:, 4], pred[:, 5], tcls))
    if world_size > 1 and rank != 0:
        return None
    stats = [np.concatenate([np.asarray(t) for t in x], 0) for x in zip(*stats)] if stats else []
    mp = mr = map50 = map_ = 0.0
    maps = np.zeros(nc)
    if len(stats) and stats[0].any():
        p, r, ap, f1, ap_class = ap_per_class(*stats, plot=False)
        ap50, ap = ap[:, 0], ap.mean(1)
        mp, mr, map50, map_ = float(p.mean()), float(r.mean()), float(ap50.mean()), float(ap.mean())
        for i, c in enumerate(ap_class):
            maps[c] = ap[i]
    return (mp, mr, map50, map_), maps
