"""CUDA-graphed detection step: Model._forward_once + fused decode/filter + batched NMS captured ONCE per
(batch, H, W, dtype) and replayed — the layer loop of models/yolo.py:211-239 costs ~150 C-ABI launches, each a ctypes
call plus a plan look-up; at detect.py's operating point (batch 1, detect.py:168-251) the GPU work of a launch is a few
microseconds, so the eager loop is host-bound there.  A replay is one `cudaGraphLaunch`.

    det = GraphedDetector(model, conf_thres=0.25, iou_thres=0.45, max_det=1000)
    dets = det(img)            # same return type as non_max_suppression(model(img)[0], ...): list of (n, 6) tensors

What makes the step capturable: every kernel of the path is launched on the current stream through the C-ABI with
caller-owned buffers (torch's graph-private pool during capture), tensor maps are kernel parameters (baked into the
graph nodes), and the NMS chain has a sync-free form with fixed-capacity candidate buffers (`ops.nms_fused_static`).
The one data-dependent size — the number of candidates — is bounded by `capacity`; the true total comes back with the
results, and a step that overflowed is repeated eagerly and the graph re-captured with larger buffers, so results are
always those of the eager path (tests compare them bit for bit).
"""
from __future__ import annotations

import torch

from . import ops
from .lazy import LazyPred
from .utils.general import Detections, non_max_suppression


class _Entry:
    __slots__ = ('graph', 'x', 'out', 'counts', 'packed', 'offsets', 'capacity', 'n')


class GraphedDetector:
    def __init__(self, model, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False, max_det=300,
                 clone_outputs=True, headroom=2.0, min_capacity=4096):
        self.model = model
        self.kw = dict(conf_thres=conf_thres, iou_thres=iou_thres, classes=classes, agnostic=agnostic, multi_label=multi_label,
                       max_det=max_det)
        self.clone_outputs = clone_outputs
        self.headroom = float(headroom)
        self.min_capacity = int(min_capacity)
        self._entries: dict = {}
        self.captures = 0
        self.eager_fallbacks = 0

    # -- capture ---------------------------------------------------------------------------------------------------
    def _eager(self, x):
        with torch.no_grad():
            pred, _ = self.model(x)
            return non_max_suppression(pred, **self.kw), pred

    def _capture(self, x, total_hint=None):
        dev = x.device
        with torch.no_grad():
            if total_hint is None:
                # warm-up on the current stream: builds every pack / plan / host-side cache the forward needs and measures
                # the candidate count that sizes the static buffers
                pred, _ = self.model(x)
                if not isinstance(pred, LazyPred):
                    raise ops.DmayError('GraphedDetector needs a model whose head is the kernel-path Detect (LazyPred output)')
                cand = ops.filter_candidates(None, self.kw['conf_thres'], multi_label=self.kw['multi_label'],
                                             classes=self.kw['classes'], levels=pred._levels, na=pred._na, nc=pred._no - 5)
                total_hint = int(cand['total'])
        n = x.shape[0]
        e = _Entry()
        e.n = n
        e.capacity = int(total_hint * self.headroom) + self.min_capacity * n
        e.x = torch.empty_like(x)
        e.x.copy_(x)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            self._step_static(e)                      # once un-captured on the capture stream (per-stream workspaces)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        e.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(e.graph, stream=side), torch.no_grad():   # same stream: per-stream workspaces are reused
            self._step_static(e)
        self.captures += 1
        return e

    def _step_static(self, e):
        pred, _ = self.model(e.x)
        kw = self.kw
        e.out, e.counts, e.packed, e.offsets = ops.nms_fused_static(
            pred._levels, pred._na, pred._no - 5, kw['conf_thres'], kw['iou_thres'], e.capacity, classes=kw['classes'],
            agnostic=kw['agnostic'], multi_label=kw['multi_label'], max_det=kw['max_det'])

    # -- replay ----------------------------------------------------------------------------------------------------
    def replay_only(self, x):
        """Copy `x` into the static input and replay; no overflow check, no host synchronisation (benchmarks)."""
        e = self._entry(x)
        e.x.copy_(x, non_blocking=True)
        e.graph.replay()
        return e

    def _entry(self, x):
        key = (tuple(x.shape), x.dtype, x.device.index)
        e = self._entries.get(key)
        if e is None:
            e = self._entries[key] = self._capture(x)
        return e

    def __call__(self, x):
        if not x.is_cuda:
            raise ops.DmayError('GraphedDetector runs on CUDA tensors only')
        e = self.replay_only(x)
        # ONE D2H carries the per-image counts and the true candidate total of this replay
        host = torch.cat((e.counts.to(torch.int64), e.offsets)).tolist()
        offs = host[e.n:]
        total = offs[-1]
        maxc = max(offs[i + 1] - offs[i] for i in range(e.n))
        # (every image reserves inside its own region of capacity // n candidates, see ops.PER_IMAGE_REGIONS)
        if total > e.capacity or maxc > e.capacity // e.n:   # the static buffers were too small: exact eager result,
            self.eager_fallbacks += 1                 # then a re-capture sized for it
            dets, _ = self._eager(x)
            key = (tuple(x.shape), x.dtype, x.device.index)
            self._entries[key] = self._capture(x, total_hint=max(total, e.n * maxc))
            return dets
        out, cnt, packed = e.out, e.counts, e.packed
        if self.clone_outputs:                        # the static buffers are overwritten by the next replay
            packed = packed.clone()
            nd = e.n * self.kw['max_det'] * 6
            out, cnt = packed[:nd].view(e.n, self.kw['max_det'], 6), packed[nd:].view(torch.int32)
        d = Detections(out, cnt, packed)
        d._pending = False
        for i, c in enumerate(host[:e.n]):
            list.__setitem__(d, i, out[i, :c])
        return d
