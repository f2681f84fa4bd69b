"""Drop-in mirror of the reference's models/detect_t.py: the anchor-free `TDetect` head with
distribution-focal-loss (DFL) box decoding, used by CASPD_ODRTA.yaml.

  per level i:  box_i = cv2[i](x_i)  (Conv3x3 -> Conv3x3 -> 1x1+bias, 4*reg_max channels)
                cls_i = cv3[i](x_i)  (Conv3x3 -> Conv3x3 -> 1x1+bias, nc channels)
  eval:         box, cls over all levels [b, 4*reg_max | nc, A];  d = DFL(box) = E_softmax(bin index)  [b, 4, A]
                (lt, rb) = d;  xywh = ((a - lt + a + rb) / 2, rb + lt) * stride,  a = cell centre (+0.5) -- detect_t.py:46-58
                y = cat(xywh, sigmoid(cls))  ->  (y, (x, box, cls))

Kernel path (CUDA, eval): the six convs of every level run on the tcgen05 conv kernel (the two 1x1 heads with fp32 output);
the DFL softmax-expectation, dist2bbox, * stride and the class sigmoid are ONE kernel per level reading the fp32 head logits
once (`dmay_dfl_decode`, csrc/post.cu) -- no cuDNN / eager arithmetic on the path.  The raw tuple `(x, box, cls)` the
reference returns next to `y` is re-laid by plain copies (torch.cat of views).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import ops
from .common import Conv, get_conv_pack, kernel_path

__all__ = ['TDetect', 'DFL', 'make_anchors', 'dist2bbox']
DFL_KERNEL = __import__('os').environ.get('DMAY_DFL_KERNEL', '1') != '0'   # A/B switch (tests compare both forms)


def make_anchors(feats, strides, grid_cell_offset=0.5):
    """Cell-centre points and per-point strides of every level, row-major (y, x) -- detect_t.py:66-79."""
    pts, strs = [], []
    dtype, device = feats[0].dtype, feats[0].device
    for f, s in zip(feats, strides):
        h, w = f.shape[-2:]
        sx = torch.arange(w, device=device, dtype=dtype) + grid_cell_offset
        sy = torch.arange(h, device=device, dtype=dtype) + grid_cell_offset
        gy, gx = torch.meshgrid(sy, sx, indexing='ij')
        pts.append(torch.stack((gx, gy), -1).view(-1, 2))
        strs.append(torch.full((h * w, 1), float(s), dtype=dtype, device=device))
    return torch.cat(pts), torch.cat(strs)


def dist2bbox(distance, anchor_points, xywh=True, dim=-1):
    """(left, top, right, bottom) distances from a point -> box -- detect_t.py:81-90."""
    lt, rb = torch.split(distance, 2, dim)
    x1y1, x2y2 = anchor_points - lt, anchor_points + rb
    if xywh:
        return torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), dim)
    return torch.cat((x1y1, x2y2), dim)


class DFL(nn.Module):
    """Expectation of the bin index under a softmax over `c1` bins, as a frozen 1x1 conv -- detect_t.py:92-102."""

    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1)
        self.c1 = c1

    def forward(self, x):
        b, _, a = x.shape
        return self.conv(x.view(b, 4, self.c1, a).transpose(2, 1).softmax(1)).view(b, 4, a)


class TDetect(nn.Module):
    """detect_t.py:23-64.  Same constructor, attributes, state_dict keys and return values."""
    shape = None
    anchors = torch.empty(0)
    strides = torch.empty(0)
    dynamic = False
    export = False

    def __init__(self, nc=80, ch=(), inplace=True):
        super().__init__()
        self.nc = nc
        self.reg_max = 16
        self.nl = len(ch)
        self.no = nc + self.reg_max * 4
        self.inplace = inplace
        self.stride = torch.zeros(self.nl)
        c2, c3 = max(ch[0] // 4, 16), max(ch[0], self.no - 4)
        self.cv2 = nn.ModuleList(
            nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for x in ch)
        self.cv3 = nn.ModuleList(
            nn.Sequential(Conv(x, c3, 3), Conv(c3, c3, 3), nn.Conv2d(c3, self.nc, 1)) for x in ch)
        self.dfl = DFL(self.reg_max)

    # -- the two branches of one level ------------------------------------------------------------------------------
    def _branch(self, seq, slot, x, fast):
        t = seq[1](seq[0](x))
        if not fast:
            return seq[2](t)
        pk = get_conv_pack(self, slot, seq[2], None, t.device)
        return ops.conv(t, pk, ops.ACT_NONE, out_fp32=True)      # fp32 logits, NHWC strides under an NCHW shape

    def forward(self, x):
        fast = kernel_path(self, x)
        x = list(x)
        shape = x[0].shape
        heads = []
        for i in range(self.nl):
            heads.append((self._branch(self.cv2[i], f'cv2.{i}.2', x[i], fast), self._branch(self.cv3[i], f'cv3.{i}.2', x[i], fast)))
            x[i] = torch.cat(heads[-1], 1)
        box, cls = torch.cat([xi.reshape(shape[0], self.no, -1) for xi in x], 2).split((self.reg_max * 4, self.nc), 1)
        if self.training:
            return x, box, cls
        if fast and DFL_KERNEL and self.dfl.c1 == self.reg_max:
            strides = self.__dict__.get('_b200_strides')
            if strides is None or strides[0] is not self.stride:
                strides = self.__dict__['_b200_strides'] = (self.stride, [float(v) for v in self.stride.tolist()])
            y = ops.dfl_decode([h[0] for h in heads], [h[1] for h in heads], self.nc, self.reg_max, strides[1])
            return y if self.export else (y, (x, box, cls))
        if self.dynamic or self.shape != shape or self.anchors.device != box.device:
            self.anchors, self.strides = (t.transpose(0, 1) for t in make_anchors(x, self.stride, 0.5))
            self.shape = shape
        dbox = dist2bbox(self.dfl(box.float()), self.anchors.unsqueeze(0).float(), xywh=True, dim=1) * self.strides.float()
        y = torch.cat((dbox, cls.float().sigmoid()), 1)
        return y if self.export else (y, (x, box, cls))

    def bias_init(self):
        """detect_t.py:60-64 -- box bias 1.0, class prior 5 objects / nc classes per 640-pixel image."""
        for a, b, s in zip(self.cv2, self.cv3, self.stride):
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[:self.nc] = math.log(5 / self.nc / (640 / s) ** 2)
