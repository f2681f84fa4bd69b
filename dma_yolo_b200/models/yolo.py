"""Drop-in mirror of the reference's models/yolo.py: Detect, Model, parse_model.

`Model(cfg, ch=3, nc=None, anchors=None).forward(x, augment=False, profile=False, visualize=False)`
returns `(pred[bs, rows, no], [x_i[bs, na, ny, nx, no]])` in eval and the list in training, exactly
as models/yolo.py:187-239 — on CUDA in eval mode every layer of the hot path runs on libdmayolo.so.
"""
from __future__ import annotations

import math
from copy import deepcopy
from pathlib import Path

import torch
import torch.nn as nn

from .. import ops
from ..lazy import LazyPred
from ..ops import ACT_NONE, Up
from ..utils.general import LOGGER, check_version, make_divisible
from ..utils.torch_utils import copy_attr, fuse_conv_and_bn, initialize_weights, model_info, scale_img, time_sync
from . import common as _common
from .common import *  # noqa: F401,F403  (module names are looked up by parse_model)
from .common import (CA, SM, AdConcat2, AdConcat3, Adapt_Add2, Adapt_Add3, Bottleneck, BottleneckCSP, C3, C3CA, C3HB, C3STR,
                     CABottleneck, Concat, Contract, CoorAttention, DWConv, Expand, Focus, SCConv, SPP, SPPCSPC, SPPF,
                     SPPFCSPC, _materialize, _PackMixin, get_conv_pack, kernel_path, space_to_depth)
from .common import GnConv, HorBlock
from .extra import (ASPP, BAM, C3SPP, C3TR, CBAM, CSPCM, AdaptADD, AdaptConcat, C3Ghost, C3GhostV2, ConvMix, CrossConv, DMConv,
                    DMMConv, DMMConv2, GhostBottleneck, GhostConv, MixConv2d, MP, SMMConv)
from .cspcm import Conv  # shadows common.Conv exactly like `from models.cspcm import *` (models/yolo.py:24)
from .detect_t import TDetect  # models/yolo.py:9

CFG_DIR = Path(__file__).resolve().parent
PAD_HEAD_ROWS = __import__('os').environ.get('DMAY_PAD_HEAD', '1') != '0'   # A/B switch: anchor rows of the head logits padded to 4n floats


def check_anchor_order(m):
    """utils/autoanchor.py:16-23"""
    a = m.anchors.prod(-1).view(-1)
    da = a[-1] - a[0]
    ds = m.stride[-1] - m.stride[0]
    if da.sign() != ds.sign():
        m.anchors[:] = m.anchors.flip(0)


class Detect(_PackMixin, nn.Module):
    """Detection head: 1x1 conv + bias per level, then grid/anchor decode — models/yolo.py:40-114."""
    stride = None
    onnx_dynamic = False

    def __init__(self, nc=80, anchors=(), ch=(), inplace=True):
        super().__init__()
        self.nc = nc
        self.no = nc + 5
        self.nl = len(anchors)
        self.na = len(anchors[0]) // 2
        self.grid = [torch.zeros(1)] * self.nl
        self.anchor_grid = [torch.zeros(1)] * self.nl
        self.register_buffer('anchors', torch.tensor(anchors).float().view(self.nl, -1, 2))
        self.m = nn.ModuleList(nn.Conv2d(x, self.no * self.na, 1) for x in ch)
        self.inplace = inplace

    def forward(self, x):
        if kernel_path(self, x) and self.na <= 5 and self.nl <= 5:
            return self.forward_b200(x)
        x = _materialize(list(x))
        z = []
        for i in range(self.nl):
            x[i] = self.m[i](x[i])
            bs, _, ny, nx = x[i].shape
            x[i] = x[i].view(bs, self.na, self.no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
            if not self.training:
                if self.onnx_dynamic or self.grid[i].shape[2:4] != x[i].shape[2:4]:
                    self.grid[i], self.anchor_grid[i] = self._make_grid(nx, ny, i)
                y = x[i].sigmoid()
                if self.inplace:
                    y[..., 0:2] = (y[..., 0:2] * 2 - 0.5 + self.grid[i]) * self.stride[i]
                    y[..., 2:4] = (y[..., 2:4] * 2) ** 2 * self.anchor_grid[i]
                else:
                    xy = (y[..., 0:2] * 2 - 0.5 + self.grid[i]) * self.stride[i]
                    wh = (y[..., 2:4] * 2) ** 2 * self.anchor_grid[i]
                    y = torch.cat((xy, wh, y[..., 4:]), -1)
                z.append(y.view(bs, -1, self.no))
        return x if self.training else (torch.cat(z, 1), x)

    def _make_grid(self, nx=20, ny=20, i=0):
        d = self.anchors[i].device
        yv, xv = torch.meshgrid([torch.arange(ny).to(d), torch.arange(nx).to(d)], indexing='ij')
        grid = torch.stack((xv, yv), 2).expand((1, self.na, ny, nx, 2)).float()
        anchor_grid = (self.anchors[i].clone() * self.stride[i]).view((1, self.na, 1, 1, 2)) \
            .expand((1, self.na, ny, nx, 2)).float()
        return grid, anchor_grid

    def _host_consts(self):
        """strides and pixel anchors as Python floats, cached per buffer version (one D2H)."""
        key = (self.anchors.data_ptr(), self.anchors._version, self.stride.data_ptr(), self.stride._version)
        c = self.__dict__.get('_b200_consts')
        if c is None or c[0] != key:
            st = self.stride.detach().float().cpu()
            an = (self.anchors.detach().float().cpu() * st.view(-1, 1, 1))  # fp32 product, as _make_grid does
            c = (key, st.tolist(), an.tolist())
            self.__dict__['_b200_consts'] = c
        return c[1], c[2]

    def _head_pack(self, i, device):
        """1x1 head conv i as a GEMM pack.  When no = 5 + nc is not a multiple of 4, every anchor's `no` output rows are
        padded to `pitch` = round_up(no, 4) with zero weight rows (the layout of the logits is ours): anchor rows of the fp32
        logits are then 16-byte aligned, which the fused decode + filter kernel needs for 16-byte staging."""
        conv = self.m[i]
        pitch = ops.round_up(self.no, 4)
        if pitch == self.no or not PAD_HEAD_ROWS:
            return get_conv_pack(self, f'm{i}', conv, None, device), 0
        key = (str(device), conv.weight.data_ptr(), conv.weight._version, conv.bias.data_ptr(), conv.bias._version, pitch)
        cache = self.__dict__.setdefault('_b200_packs', {})
        pk = cache.get(f'mp{i}')
        if pk is None or pk.key != key:
            w = conv.weight.detach().float()
            cin = w.shape[1]
            wp = torch.zeros((self.na, pitch, cin, 1, 1), dtype=torch.float32, device=w.device)
            wp[:, :self.no] = w.view(self.na, self.no, cin, 1, 1)
            bp = torch.zeros((self.na, pitch), dtype=torch.float32, device=w.device)
            bp[:, :self.no] = conv.bias.detach().float().view(self.na, self.no)
            pk = ops.pack_conv(wp.view(self.na * pitch, cin, 1, 1), conv_bias=bp.view(-1), device=device)
            pk.key = key
            cache[f'mp{i}'] = pk
        return pk, pitch

    def forward_b200(self, x):
        """Raw fp32 head logits (NHWC) from the tcgen05 GEMM; decode is deferred to LazyPred /
        non_max_suppression so that it can be fused with the confidence filter."""
        strides, anchors_px = self._host_consts()
        levels, xs = [], []
        for i in range(self.nl):
            t = _materialize(x[i])
            pk, pitch = self._head_pack(i, t.device)
            # two N tiles of 160 columns cover a padded 3 x 88 = 264-channel head with 21 % spare columns (256 + 16 would
            # spend a whole second tile on 8 channels)
            bn = 160 if (pitch and 256 < pk.cout_pad <= 320) else 0
            lg = ops.conv(t, pk, ACT_NONE, out_fp32=True, block_n=bn)   # [bs, na*pitch, ny, nx] view of an NHWC slab
            bs, _, ny, nx = lg.shape
            ld = ops.ld_of(lg)
            nhwc = lg.permute(0, 2, 3, 1)                          # [bs, ny, nx, na*pitch], strides (.., ld, 1)
            levels.append(ops.DetectLevel(logits=lg, stride=strides[i], anchors_px=[tuple(a) for a in anchors_px[i]],
                                          ny=ny, nx=nx, ld=ld, pitch=pitch))
            v = nhwc.unflatten(-1, (self.na, pitch or self.no))
            xs.append((v[..., :self.no] if pitch else v).permute(0, 3, 1, 2, 4))   # (bs,na,ny,nx,no) view
        return LazyPred(levels, self.na, self.no), xs

    def __getstate__(self):
        d = _PackMixin.__getstate__(self)
        d.pop('_b200_consts', None)
        return d


class Model(nn.Module):
    """YOLOv5-style model built from a YAML/dict config — models/yolo.py:117-350."""

    def __init__(self, cfg='yolov5s.yaml', ch=3, nc=None, anchors=None):
        super().__init__()
        if isinstance(cfg, dict):
            self.yaml = cfg
        else:
            import yaml
            p = Path(cfg)
            if not p.exists():   # a bare config name (optionally hub/<name>) resolves against the packaged layer tables
                for cand in (CFG_DIR / cfg, CFG_DIR / p.name):
                    if cand.exists():
                        p = cand
                        break
            self.yaml_file = p.name
            with open(p, errors='ignore') as f:
                self.yaml = yaml.safe_load(f)
        ch = self.yaml['ch'] = self.yaml.get('ch', ch)
        if nc and nc != self.yaml['nc']:
            LOGGER.info(f"Overriding model.yaml nc={self.yaml['nc']} with nc={nc}")
            self.yaml['nc'] = nc
        if anchors:
            LOGGER.info(f'Overriding model.yaml anchors with anchors={anchors}')
            self.yaml['anchors'] = round(anchors)
        self.model, self.save = parse_model(deepcopy(self.yaml), ch=[ch])
        self.names = [str(i) for i in range(self.yaml['nc'])]
        self.inplace = self.yaml.get('inplace', True)

        m = self.model[-1]
        if isinstance(m, Detect):
            s = 256  # 2x min stride
            m.inplace = self.inplace
            m.stride = torch.tensor([s / x.shape[-2] for x in self.forward(torch.zeros(1, ch, s, s))])
            m.anchors /= m.stride.view(-1, 1, 1)
            check_anchor_order(m)
            self.stride = m.stride
            self._initialize_biases()
        elif isinstance(m, TDetect):   # models/yolo.py:173-180 — anchor-free head: strides from the train-mode probe, then bias_init
            s = 256
            m.inplace = self.inplace
            m.stride = torch.tensor([s / x.shape[-2] for x in self.forward(torch.zeros(1, ch, s, s))[0]])
            self.stride = m.stride
            m.bias_init()
        initialize_weights(self)

    def forward(self, x, augment=False, profile=False, visualize=False):
        if augment:
            return self._forward_augment(x)
        return self._forward_once(x, profile, visualize)

    def _forward_augment(self, x):
        img_size = x.shape[-2:]
        s = [1, 1, 0.83, 0.83, 0.67, 0.67]
        f = [None, 3, None, 3, None, 3]
        y = []
        for si, fi in zip(s, f):
            xi = scale_img(x.flip(fi) if fi else x, si, gs=int(self.stride.max()))
            yi = self._forward_once(xi)[0]
            yi = self._descale_pred(yi + 0, fi, si, img_size)
            y.append(yi)
        y = self._clip_augmented(y)
        return torch.cat(y, 1), None

    def _forward_once(self, x, profile=False, visualize=False):
        """Layer loop with `m.f` routing — models/yolo.py:211-239.  In the CUDA-eval path a nearest
        `nn.Upsample(2^k)` is not materialised when all its consumers can read the low-resolution source
        (AdConcat / Concat fuse it), otherwise it runs on the upsample kernel."""
        y, dt = [], []
        fast = kernel_path(self, x)
        for m in self.model:
            if m.f != -1:
                x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
            if profile:
                self._profile_one_layer(m, x, dt)
            nxt = self.model[m.i + 1] if fast and m.i + 1 < len(self.model) else None
            if (nxt is not None and isinstance(nxt, _common.SCConv) and nxt.f == -1 and not isinstance(x, (list, tuple, Up))
                    and getattr(type(m), 'forward_b200', None) is _common.Conv.forward_b200):
                x = m.forward_b200(x, pool4=True)      # the SCConv's AvgPool2d(4) input comes out of this conv's epilogue
            else:
                x = self._run_layer(m, x, fast)
            y.append(x if m.i in self.save else None)
            tr = self.__dict__.get('_trace')
            if tr is not None:  # test/debug hook: per-layer outputs (incl. layers that bypass nn.Module.__call__)
                tr.append(x)
        return x

    def _run_layer(self, m, x, fast=None):
        """One layer of the loop: kernel path where the module has one, torch body on CUDA otherwise."""
        if fast is None:
            fast = kernel_path(self, x)
        if fast and isinstance(m, nn.Upsample):
            return self._upsample_b200(m, x)
        fb = getattr(type(m), 'forward_b200', None)
        cv12 = (_common.C3.forward_b200, _common.C3STR.forward_b200, _common.C3HB.forward_b200)   # start with C3._cv12_slab
        if fast and isinstance(x, ops.VCat) and fb in cv12 + (_common.Conv.forward_b200,):
            return m(x)      # a 1x1 consumer reads the parts of the concat in place (ops.VCat)
        if fast and isinstance(x, ops.SPDView) and fb in cv12:
            return m(x)      # space_to_depth + cv1 | cv2 as one 2x2 / stride-2 GEMM (ops.SPDView)
        if fast and not isinstance(m, (_common.AdConcat2, _common.AdConcat3, _common.Concat)):
            x = _materialize(x)
        if fast and not _has_kernel_path(m):
            return _common.torch_body(m, m, x)
        return m(x)

    @staticmethod
    def _upsample_b200(m, x):
        x = _materialize(x)
        sf = m.scale_factor
        sf = sf if isinstance(sf, (int, float)) else (sf[0] if sf is not None and sf[0] == sf[1] else None)
        if m.mode == 'nearest' and m.size is None and sf is not None and float(sf) == int(sf) and int(sf) >= 1:
            f = int(sf)
            if f == 1:
                return x
            if f & (f - 1) == 0 and x.dim() == 4 and x.shape[1] % 8 == 0:
                return Up(ops.as_act(x), f.bit_length() - 1)
            return ops.upsample(x, f)
        return _common.torch_body(m, m, x)

    def _profile_one_layer(self, m, x, dt):
        c = isinstance(m, (Detect, TDetect))
        t = time_sync()
        for _ in range(10):
            m(x.copy() if c else x)
        dt.append((time_sync() - t) * 100)
        LOGGER.info(f'{dt[-1]:10.2f} {0:10.2f} {m.np:10.0f}  {m.type}')

    def _descale_pred(self, p, flips, scale, img_size):
        if self.inplace:
            p[..., :4] /= scale
            if flips == 2:
                p[..., 1] = img_size[0] - p[..., 1]
            elif flips == 3:
                p[..., 0] = img_size[1] - p[..., 0]
        else:
            x, y, wh = p[..., 0:1] / scale, p[..., 1:2] / scale, p[..., 2:4] / scale
            if flips == 2:
                y = img_size[0] - y
            elif flips == 3:
                x = img_size[1] - x
            p = torch.cat((x, y, wh, p[..., 4:]), -1)
        return p

    def _clip_augmented(self, y):
        nl = self.model[-1].nl
        g = sum(4 ** x for x in range(nl))
        e = 1
        i = (y[0].shape[1] // g) * sum(4 ** x for x in range(e))
        y[0] = y[0][:, :-i]
        i = (y[-1].shape[1] // g) * sum(4 ** (nl - 1 - x) for x in range(e))
        y[-1] = y[-1][:, i:]
        return y

    def _initialize_biases(self, cf=None):
        """models/yolo.py:293-301 — obj/cls bias priors."""
        m = self.model[-1]
        for mi, s in zip(m.m, m.stride):
            b = mi.bias.view(m.na, -1)
            b.data[:, 4] += math.log(8 / (640 / s) ** 2)
            b.data[:, 5:] += math.log(0.6 / (m.nc - 0.999999)) if cf is None else torch.log(cf / cf.sum())
            mi.bias = torch.nn.Parameter(b.view(-1), requires_grad=True)

    def fuse(self):
        """models/yolo.py:315-323 — fold BN into the conv of `Conv`/`DWConv` modules.  (The kernel path
        folds BN for every conv+BN pair regardless; this keeps the reference's API and its CPU behaviour.)"""
        for m in self.model.modules():
            if isinstance(m, (Conv, DWConv)) and hasattr(m, 'bn'):
                m.conv = fuse_conv_and_bn(m.conv, m.bn)
                delattr(m, 'bn')
                m.forward = m.forward_fuse
        return self

    def info(self, verbose=False, img_size=640):
        return model_info(self, verbose, img_size)

    def _apply(self, fn):
        self = super()._apply(fn)
        m = self.model[-1]
        if isinstance(m, Detect):
            m.stride = fn(m.stride)
            m.grid = list(map(fn, m.grid))
            if isinstance(m.anchor_grid, list):
                m.anchor_grid = list(map(fn, m.anchor_grid))
        elif isinstance(m, TDetect):   # models/yolo.py:345-348
            m.stride = fn(m.stride)
            m.anchors = fn(m.anchors)
            m.strides = fn(m.strides)
        return self


def _has_kernel_path(m: nn.Module) -> bool:
    if isinstance(m, nn.Sequential):
        return len(m) > 0 and all(_has_kernel_path(c) for c in m)
    return type(m).__module__.startswith(__name__.rsplit('.', 1)[0])


_CHANNEL_AWARE = None


def parse_model(d, ch):
    """YAML dict -> nn.Sequential + save list — models/yolo.py:353-478 (channel bookkeeping per module type)."""
    anchors, nc, gd, gw = d['anchors'], d['nc'], d['depth_multiple'], d['width_multiple']
    na = (len(anchors[0]) // 2) if isinstance(anchors, list) else anchors
    no = na * (nc + 5)
    # models/yolo.py:387-388 (the channel-aware list) and 399 (modules that take the repeat count as an argument)
    scaled = [Conv, _common.Conv, GhostConv, Bottleneck, GhostBottleneck, SPP, SPPF, DWConv, MixConv2d, Focus, CrossConv,
              BottleneckCSP, C3, C3TR, C3STR, C3SPP, C3Ghost, ASPP, CBAM, nn.ConvTranspose2d, CoorAttention, CABottleneck, C3CA,
              SPPCSPC, SPPFCSPC, SCConv, HorBlock, C3HB, GnConv]
    repeated = [BottleneckCSP, C3, C3TR, C3STR, C3Ghost, C3CA, C3HB, BAM]
    layers, save, c2 = [], [], ch[-1]
    for i, (f, n, m, args) in enumerate(d['backbone'] + d['head']):
        if isinstance(m, str):
            try:
                m = eval(m)
            except NameError as e:
                raise NotImplementedError(f"layer {i}: module '{m}' is outside the accelerated detection path "
                                          f"(SURVEY.md 8f lists it as a later row)") from e
        for j, a in enumerate(args):
            try:
                args[j] = eval(a) if isinstance(a, str) else a
            except NameError:
                pass
        n = n_ = max(round(n * gd), 1) if n > 1 else n
        if m in scaled:
            c1, c2 = ch[f], args[0]
            if c2 != no:
                c2 = make_divisible(c2 * gw, 8)
            args = [c1, c2, *args[1:]]
            if m in repeated:
                args.insert(2, n)
                n = 1
        elif m is nn.BatchNorm2d:
            args = [ch[f]]
        elif m in (Concat, AdConcat2, AdConcat3):
            c2 = sum(ch[x] for x in f)
        elif m in (ConvMix, CSPCM):                       # models/yolo.py:410-414
            c1, c2 = ch[f], args[0]
            if c2 != no:
                c2 = make_divisible(c2 * gw, 8)
            args = [c1, c2, *args[1:]]
        elif m in (AdaptConcat, AdaptADD):                # models/yolo.py:416-420
            c2 = sum(ch[x] for x in f)
            args = [len(f), *args]
        elif m in (Adapt_Add2, Adapt_Add3):
            c2 = max([ch[x] for x in f])
        elif m is C3GhostV2:                              # models/yolo.py:424-431
            c1, c2 = ch[f], args[0]
            if c2 != no:
                c2 = make_divisible(c2 * gw, 8)
            args = [c1, c2, *args[1:]]
            args.insert(2, n)
            n = 1
        elif m is Detect:
            args.append([ch[x] for x in f])
            if isinstance(args[1], int):
                args[1] = [list(range(args[1] * 2))] * len(f)
        elif m is TDetect:   # models/yolo.py:438-439
            args.append([ch[x] for x in f])
        elif m is Contract:
            c2 = ch[f] * args[0] ** 2
        elif m is Expand:
            c2 = ch[f] // args[0] ** 2
        elif m is space_to_depth:                        # models/yolo.py:446-447 (`SM` is not listed there: c2 = ch[f])
            c2 = 4 * ch[f]
        elif m is SMMConv:                                # models/yolo.py:448-463
            c1, c2 = ch[f], 4 * args[0]
            args = [c1, args[0]]
        elif m is DMMConv:
            c1, c2 = ch[f], 5 * args[0]
            args = [c1, args[0]]
        elif m is DMMConv2:
            c1 = ch[f]
            c2 = args[0] + 4 * c1
            args = [c1, args[0]]
        elif m is DMConv:
            c1, c2 = ch[f], 4 * args[0]
            args = [c1, args[0]]
        else:
            c2 = ch[f]
        m_ = nn.Sequential(*(m(*args) for _ in range(n))) if n > 1 else m(*args)
        t = str(m)[8:-2].replace('__main__.', '')
        np_ = sum(x.numel() for x in m_.parameters())
        m_.i, m_.f, m_.type, m_.np = i, f, t, np_
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
        layers.append(m_)
        if i == 0:
            ch = []
        ch.append(c2)
    return nn.Sequential(*layers), sorted(save)
