"""val.py-style evaluation on the B200 path: forward -> fused decode/filter -> batched NMS -> scale_coords + IoU matching
(`process_batch`) on the GPU, ONE device-to-host copy per batch (padded detections, counts and the `correct` matrix),
the detections of all ranks all-gathered first (NCCL, one collective), then the reference's host-side statistics
(`ap_per_class`, numpy) unchanged.

`process_batch` and the statistics block follow val.py:62-83 and val.py:236-288; `run` accepts the in-memory
dataloader form that the reference's `val.run(model=..., dataloader=...)` call takes (the file-based data pipeline is
out of scope).  On CPU tensors (no CUDA) the reference's host body runs instead of the kernel."""
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


def _geometry(img1_shape, img0_shape, ratio_pad):
    """(gain, pad_x, pad_y) of scale_coords (utils/general.py:605-612), computed in Python doubles like the reference."""
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    return float(gain), float(pad[0]), float(pad[1])


def match_batch_device(padded, counts, targets, shapes, img_hw, iouv, single_cls=False, first_image=0):
    """`process_batch` of every image of a batch in one kernel (libdmayolo `dmay_val_match`).
    padded [B, max_det, 6] / counts [B]: the NMS output on the device.  targets [n, 6] on the host = (image index, cls,
    xywh in network-input pixels), as val.py:227 leaves them.  The label boxes are converted and rescaled on the host with the
    reference's own fp32 arithmetic (a few hundred values), uploaded in one small copy together with the per-image geometry.
    -> (correct u8 [B, max_det, niou] on the device, per-image lists of target classes)."""
    from . import ops
    B, max_det = padded.shape[0], padded.shape[1]
    height, width = img_hw
    labs, offs, geom, tcls = [], [0], [], []
    for si in range(B):
        labels = targets[targets[:, 0] == si + first_image, 1:]
        shape = shapes[si][0] if shapes is not None else (height, width)
        ratio_pad = shapes[si][1] if shapes is not None else None
        tcls.append(labels[:, 0].tolist() if len(labels) else [])
        if len(labels):
            tbox = xywh2xyxy(labels[:, 1:5])
            scale_coords((height, width), tbox, shape, ratio_pad)        # labels in original-image pixels (val.py:262-264)
            labs.append(torch.cat((labels[:, 0:1], tbox), 1).float())
        offs.append(offs[-1] + len(labels))
        g, px, py = _geometry((height, width), shape, ratio_pad)
        geom.append([g, px, py, float(shape[0]), float(shape[1])])
    lab = torch.cat(labs, 0) if labs else torch.zeros((0, 5))
    max_labels = max((offs[i + 1] - offs[i] for i in range(B)), default=0)
    dev = padded.device
    # one small H2D for the batch's labels and geometry
    host = torch.cat((lab.reshape(-1), torch.tensor(geom, dtype=torch.float32).reshape(-1), iouv.float().reshape(-1)))
    devbuf = host.to(dev, non_blocking=True)
    nl5, ng = lab.numel(), 5 * B
    lab_d = devbuf[:nl5] if nl5 else torch.zeros(5, device=dev)
    geom_d, iouv_d = devbuf[nl5:nl5 + ng], devbuf[nl5 + ng:]
    off_d = torch.tensor(offs, dtype=torch.int32).to(dev, non_blocking=True)
    correct = torch.empty((B, max_det, iouv.numel()), dtype=torch.uint8, device=dev)
    ops.call('dmay_val_match', torch.cuda.current_stream(dev).cuda_stream, det=padded.data_ptr(), counts=counts.data_ptr(),
             labels=lab_d.data_ptr(), lab_off=off_d.data_ptr(), geom=geom_d.data_ptr(), iouv=iouv_d.data_ptr(),
             correct=correct.data_ptr(), B=B, max_det=max_det, niou=int(iouv.numel()), max_labels=int(max_labels),
             single_cls=int(bool(single_cls)))
    return correct, tcls


@torch.no_grad()
def run(data, model=None, dataloader=None, batch_size=32, imgsz=640, conf_thres=0.001, iou_thres=0.6, single_cls=False,
        half=False, plots=False, rank=0, world_size=1, **_ignored):
    """-> ((mp, mr, map50, map), maps[nc]) like the head of the reference's return value (val.py:333-347).
    `dataloader` yields (img uint8/float [B,3,H,W], targets [n,6] = (img_idx, cls, xywh normalised), paths, shapes).
    With world_size > 1 every rank passes ITS images (img, shapes of its slice) and the batch's GLOBAL targets / shapes are
    only needed on rank 0 (image index = rank * B + local index: rank-then-image order); rank 0 returns the statistics."""
    from .dist import all_gather_detections, pad_detections
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
        if getattr(out, 'packed', None) is not None:
            padded, counts, packed = out.padded, out.counts, out.packed
        else:                                # CPU body of non_max_suppression: a plain list
            padded, counts = pad_detections(out, max_det, device)
            packed = None
        if world_size > 1:   # rank-then-image order (np.argsort in ap_per_class is order sensitive)
            padded, counts = all_gather_detections(padded, counts, packed=packed)
            if rank != 0:
                continue
        nimg = padded.shape[0]
        all_shapes = shapes if (shapes is None or len(shapes) == nimg) else None
        if padded.is_cuda:
            # scale_coords + IoU matching of the whole batch on the device; ONE D2H carries detections, counts and `correct`
            correct_d, tcls_all = match_batch_device(padded, counts, targets, all_shapes, (height, width), iouv, single_cls)
            det_h, cnt_h, cor_h = padded.cpu(), counts.tolist(), correct_d.cpu().bool()
            for si in range(nimg):
                n, tcls = cnt_h[si], tcls_all[si]
                if n == 0:
                    if len(tcls):
                        stats.append((torch.zeros(0, iouv.numel(), dtype=torch.bool), torch.Tensor(), torch.Tensor(), tcls))
                    continue
                pred = det_h[si, :n]
                cls = torch.zeros(n) if single_cls else pred[:, 5]
                correct = cor_h[si, :n] if len(tcls) else torch.zeros(n, iouv.numel(), dtype=torch.bool)
                stats.append((correct, pred[:, 4], cls, tcls))
            continue
        cnt_h = counts.tolist()
        for si in range(nimg):               # host body (CPU tensors): val.py:239-274 as is
            pred = padded[si, :cnt_h[si]].float().cpu()
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
            shape = all_shapes[si][0] if all_shapes is not None else (height, width)
            ratio_pad = all_shapes[si][1] if all_shapes is not None else None
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
