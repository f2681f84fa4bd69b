"""YAML module names OUTSIDE the accelerated path ("run as reference torch ops on CUDA").

Same class names, constructor signatures, submodule names and creation order as the reference (state_dict keys and the
seeded initialisation stay identical — tests/test_modules_cpu.py checks the digests of every reference YAML), so that
every `models/*.yaml` the reference builds also builds here.  None of these classes has a kernel of its own: on the
CUDA eval path the whole subtree runs its torch-op body in the parameters' dtype (`_Ref.forward`); the layers around
them stay on the kernels, activations are converted at the boundary (`ops.as_act`).

References: models/common.py:184-210 (C3TR, C3SPP, C3Ghost), 229-241 (ASPP), 260-355 (CBAM, TransformerLayer/Block),
666-699 (GhostConv, GhostBottleneck), 913-992 (AdaptADD, AdaptConcat), 1063-1081 (add_conv), 1441-1600 (C3GhostV2, MP,
SMMConv, DMMConv2, DMMConv, DMConv, DMMixConv2d, BAM); models/cspcm.py:25-54 (ConvMix, CSPCM); models/GhostV2.py;
models/experimental.py:15-90 (CrossConv, MixConv2d).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import common as C
from .common import C3, SM, SPP, CABottleneck, Conv, DWConv


class _Ref(nn.Module):
    """forward = the reference torch-op body (`_ref`); on the CUDA eval path the subtree is pinned to torch ops."""

    def forward(self, x):
        if C.kernel_path(self, x):
            with C.reference_ops():
                return C.torch_body(self, self._ref, x)
        return self._ref(x)


# ---- models/common.py -------------------------------------------------------------------------------------------------
class TransformerLayer(_Ref):
    def __init__(self, c, num_heads):
        super().__init__()
        self.ln1 = nn.LayerNorm(c)
        self.q = nn.Linear(c, c, bias=False)
        self.k = nn.Linear(c, c, bias=False)
        self.v = nn.Linear(c, c, bias=False)
        self.ma = nn.MultiheadAttention(embed_dim=c, num_heads=num_heads)
        self.ln2 = nn.LayerNorm(c)
        self.fc1 = nn.Linear(c, 4 * c, bias=False)
        self.fc2 = nn.Linear(4 * c, c, bias=False)
        self.dropout = nn.Dropout(0.1)
        self.act = nn.ReLU(True)

    def _ref(self, x):
        y = self.ln1(x)
        x = self.dropout(self.ma(self.q(y), self.k(y), self.v(y))[0]) + x
        y = self.fc2(self.dropout(self.act(self.fc1(self.ln2(x)))))
        return x + self.dropout(y)


class TransformerBlock(_Ref):
    def __init__(self, c1, c2, num_heads, num_layers):
        super().__init__()
        self.conv = None
        if c1 != c2:
            self.conv = Conv(c1, c2)
        self.linear = nn.Linear(c2, c2)
        self.tr = nn.Sequential(*(TransformerLayer(c2, num_heads) for _ in range(num_layers)))
        self.c2 = c2

    def _ref(self, x):
        if self.conv is not None:
            x = self.conv(x)
        b, _, w, h = x.shape
        p = x.flatten(2).unsqueeze(0).transpose(0, 3).squeeze(3)
        return self.tr(p + self.linear(p)).unsqueeze(3).transpose(0, 3).reshape(b, self.c2, w, h)


class C3TR(C3):
    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__(c1, c2, n, shortcut, g, e)
        c_ = int(c2 * e)
        self.m = TransformerBlock(c_, c_, 4, n)


class C3SPP(C3):
    def __init__(self, c1, c2, k=(5, 9, 13), n=1, shortcut=True, g=1, e=0.5):
        super().__init__(c1, c2, n, shortcut, g, e)
        c_ = int(c2 * e)
        self.m = SPP(c_, c_, k)


class GhostConv(_Ref):
    def __init__(self, c1, c2, k=1, s=1, g=1, act=True):
        super().__init__()
        c_ = c2 // 2
        self.cv1 = Conv(c1, c_, k, s, None, g, act)
        self.cv2 = Conv(c_, c_, 5, 1, None, c_, act)

    def _ref(self, x):
        y = self.cv1(x)
        return torch.cat([y, self.cv2(y)], 1)


class GhostBottleneck(_Ref):
    def __init__(self, c1, c2, k=3, s=1):
        super().__init__()
        c_ = c2 // 2
        self.conv = nn.Sequential(GhostConv(c1, c_, 1, 1),
                                  DWConv(c_, c_, k, s, act=False) if s == 2 else nn.Identity(),
                                  GhostConv(c_, c2, 1, 1, act=False))
        self.shortcut = nn.Sequential(DWConv(c1, c1, k, s, act=False), Conv(c1, c2, 1, 1, act=False)) if s == 2 else nn.Identity()

    def _ref(self, x):
        return self.conv(x) + self.shortcut(x)


class C3Ghost(C3):
    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__(c1, c2, n, shortcut, g, e)
        c_ = int(c2 * e)
        self.m = nn.Sequential(*(GhostBottleneck(c_, c_) for _ in range(n)))


class ASPP(_Ref):
    def __init__(self, c1, c2, k=(5, 9, 13)):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=1, padding=1)
        self.m = nn.ModuleList([nn.Conv2d(c_, c_, kernel_size=3, stride=1, padding=(x - 1) // 2, dilation=(x - 1) // 2, bias=False)
                                for x in k])
        self.cv2 = Conv(c_ * (len(k) + 2), c2, 1, 1)

    def _ref(self, x):
        x = self.cv1(x)
        return self.cv2(torch.cat([x, self.maxpool(x)] + [m(x) for m in self.m], 1))


class ChannelAttentionModule(_Ref):
    def __init__(self, c1, reduction=16):
        super().__init__()
        mid = c1 // reduction
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.max_pool = nn.AdaptiveMaxPool2d(1)
        self.shared_MLP = nn.Sequential(nn.Linear(in_features=c1, out_features=mid), nn.ReLU(),
                                        nn.Linear(in_features=mid, out_features=c1))
        self.sigmoid = nn.Sigmoid()

    def _ref(self, x):
        n = x.size(0)
        a = self.shared_MLP(self.avg_pool(x).view(n, -1))
        m = self.shared_MLP(self.max_pool(x).view(n, -1))
        return self.sigmoid(a + m)[:, :, None, None]


class SpatialAttentionModule(_Ref):
    def __init__(self):
        super().__init__()
        self.conv2d = nn.Conv2d(in_channels=2, out_channels=1, kernel_size=7, stride=1, padding=3)
        self.sigmoid = nn.Sigmoid()

    def _ref(self, x):
        s = torch.cat([torch.mean(x, dim=1, keepdim=True), torch.max(x, dim=1, keepdim=True)[0]], dim=1)
        return self.sigmoid(self.conv2d(s))


class CBAM(_Ref):
    def __init__(self, c1, c2):
        super().__init__()
        self.channel_attention = ChannelAttentionModule(c1)
        self.spatial_attention = SpatialAttentionModule()

    def _ref(self, x):
        out = self.channel_attention(x) * x
        return self.spatial_attention(out) * out


def add_conv(in_ch, out_ch, ksize=1, stride=1):
    """conv (bias-free) + BatchNorm + LeakyReLU(0.1) as a named Sequential — models/common.py:1063-1081."""
    stage = nn.Sequential()
    stage.add_module('conv', nn.Conv2d(in_channels=in_ch, out_channels=out_ch, kernel_size=ksize, stride=stride,
                                       padding=(ksize - 1) // 2, bias=False))
    stage.add_module('batch_norm', nn.BatchNorm2d(out_ch))
    stage.add_module('leaky', nn.LeakyReLU(0.1))
    return stage


class AdaptADD(_Ref):
    def __init__(self, level, out_ch, dimension, dim1, dim2, dim3=1, rfb=False):
        super().__init__()
        self.level = level
        self.d = dimension
        self.dims = [dim1, dim2, dim3]
        cc = 8 if rfb else 16
        self.compress_level = add_conv(self.dims[2], self.dims[0], 1, 1)
        self.weight_map = add_conv(self.dims[0], cc, 1, 1)
        self.weight_levels = nn.Conv2d(cc * level, level, kernel_size=1, stride=1, padding=0)
        self.expand = add_conv(self.dims[0], out_ch, 3, 1)

    def _ref(self, x):
        feats = [x[0], x[1]]
        if self.level == 3:
            feats.append(self.compress_level(x[2]))
        lw = F.softmax(self.weight_levels(torch.cat([self.weight_map(t) for t in feats], self.d)), dim=1)
        fused = sum(t * lw[:, i:i + 1] for i, t in enumerate(feats))
        return self.expand(fused)


class AdaptConcat(_Ref):
    def __init__(self, level, dimension, dim1, dim2, dim3=1, rfb=False):
        super().__init__()
        self.level = level
        self.d = dimension
        self.dims = [dim1, dim2, dim3]
        cc = 8 if rfb else 16
        self.weight_map0 = add_conv(self.dims[0], cc, 1, 1)
        self.weight_map1 = add_conv(self.dims[1], cc, 1, 1)
        self.weight_map2 = add_conv(self.dims[2], cc, 1, 1)
        self.weight_levels = nn.Conv2d(cc * level, level, kernel_size=1, stride=1, padding=0)

    def _ref(self, x):
        maps = [self.weight_map0(x[0]), self.weight_map1(x[1])]
        if self.level == 3:
            maps.append(self.weight_map2(x[2]))
        lw = F.softmax(self.weight_levels(torch.cat(maps, self.d)), dim=1)
        return torch.cat([x[i] * lw[:, i:i + 1] for i in range(len(maps))], 1)


class MP(_Ref):
    def __init__(self, k=2):
        super().__init__()
        self.m = nn.MaxPool2d(kernel_size=k, stride=k)

    def _ref(self, x):
        return self.m(x)


class SMMConv(_Ref):
    def __init__(self, c1, c2):
        super().__init__()
        c_ = int(c1 / 2)
        self.cv1 = Conv(c1, c_, 3, 1)
        self.cv2 = Conv(c1, c_, 5, 1)
        self.sm = SM()

    def _ref(self, x):
        return self.sm(torch.cat([self.cv1(x), self.cv2(x)], 1))


class DMMConv2(_Ref):
    def __init__(self, c1, c2):
        super().__init__()
        self.cv1 = Conv(c1, c2, 1, 1)
        self.sm = SM()
        self.mp = MP()

    def _ref(self, x):
        return torch.cat([self.sm(x), self.cv1(self.mp(x))], 1)


class DMMConv(_Ref):
    def __init__(self, c1, c2):
        super().__init__()
        self.cv1 = Conv(c1, c2, 1, 1)
        self.cv2 = Conv(c1, c2, 3, 1)
        self.sm = SM()
        self.mp = MP()

    def _ref(self, x):
        return torch.cat([self.sm(self.cv2(x)), self.cv1(self.mp(x))], 1)


class DMConv(_Ref):
    def __init__(self, c1, c2):
        super().__init__()
        self.cv1 = Conv(c1, c2, 3, 1)
        self.sm = SM()

    def _ref(self, x):
        return self.sm(self.cv1(x))


def _mix_channels(c2, k, equal_ch):
    """Channel split of MixConv2d / DMMixConv2d (models/experimental.py:71-81)."""
    n = len(k)
    if equal_ch:
        i = torch.linspace(0, n - 1E-6, c2).floor()
        return [int((i == g).sum()) for g in range(n)]
    b = [c2] + [0] * n
    a = np.eye(n + 1, n, k=-1)
    a -= np.roll(a, 1, axis=1)
    a *= np.array(k) ** 2
    a[0] = 1
    return [int(v) for v in np.linalg.lstsq(a, b, rcond=None)[0].round()]


class MixConv2d(_Ref):
    def __init__(self, c1, c2, k=(1, 3), s=1, equal_ch=True):
        super().__init__()
        c_ = _mix_channels(c2, k, equal_ch)
        self.m = nn.ModuleList([nn.Conv2d(c1, cc, kk, s, kk // 2, groups=math.gcd(c1, cc), bias=False) for kk, cc in zip(k, c_)])
        self.bn = nn.BatchNorm2d(c2)
        self.act = nn.SiLU()

    def _ref(self, x):
        return self.act(self.bn(torch.cat([m(x) for m in self.m], 1)))


class DMMixConv2d(MixConv2d):
    pass


class BAM(C3):
    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__(c1, c2, n, shortcut, g, e)
        c_ = int(c2 * e)
        self.m = nn.Sequential(*(CABottleneck(c_, c_, shortcut, g, e=1.0) for _ in range(n)))


# ---- models/experimental.py --------------------------------------------------------------------------------------------
class CrossConv(_Ref):
    def __init__(self, c1, c2, k=3, s=1, g=1, e=1.0, shortcut=False):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, (1, k), (1, s))
        self.cv2 = Conv(c_, c2, (k, 1), (s, 1), g=g)
        self.add = shortcut and c1 == c2

    def _ref(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


# ---- models/cspcm.py ---------------------------------------------------------------------------------------------------
class ConvMix(_Ref):
    def __init__(self, dim, dim1, kernel_size=9):
        super().__init__()
        self.Resnet = nn.Sequential(nn.Conv2d(dim, dim, kernel_size=kernel_size, groups=dim, padding='same'), nn.GELU(),
                                    nn.BatchNorm2d(dim))
        self.Conv_1x1 = nn.Sequential(nn.Conv2d(dim, dim, kernel_size=1), nn.GELU(), nn.BatchNorm2d(dim))

    def _ref(self, x):
        return self.Conv_1x1(x + self.Resnet(x))


class CSPCM(_Ref):
    def __init__(self, c1, c2, n=1, e=0.5):
        super().__init__()
        from .cspcm import Conv as CConv   # models/cspcm.py builds these from ITS Conv (pickled class path)
        c_ = int(c2 * e)
        self.cv1 = CConv(c1, c_, 1, 1)
        self.cv2 = CConv(c1, c_, 1, 1)
        self.cv3 = CConv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(ConvMix(c_, c_) for _ in range(n)))

    def _ref(self, x):
        return self.cv3(torch.cat((self.m(self.cv1(x)), self.cv2(x)), dim=1))


# ---- models/GhostV2.py -------------------------------------------------------------------------------------------------
class MyHSigmoid(nn.Module):
    def __init__(self):
        super().__init__()
        self.relu6 = nn.ReLU6()

    def forward(self, x):
        return self.relu6(x + 3.) * 0.16666667


class Activation(nn.Module):
    def __init__(self, act_func):
        super().__init__()
        table = {'relu': nn.ReLU, 'relu6': nn.ReLU6, 'sigmoid': nn.Sigmoid, 'hsigmoid': MyHSigmoid, 'hard_sigmoid': MyHSigmoid,
                 'hswish': nn.Hardswish, 'hard_swish': nn.Hardswish}
        if act_func not in table:
            raise NotImplementedError
        self.act = table[act_func]()

    def forward(self, x):
        return self.act(x)


class GlobalAvgPooling(nn.Module):
    def __init__(self):
        super().__init__()
        self.mean = nn.AdaptiveAvgPool2d(1)

    def forward(self, x):
        return self.mean(x)


class SE(nn.Module):
    def __init__(self, num_out, ratio=4):
        super().__init__()
        num_mid = int(np.ceil((num_out // ratio) * 1. / 4) * 4)
        self.pool = GlobalAvgPooling()
        self.conv_reduce = nn.Conv2d(in_channels=num_out, out_channels=num_mid, kernel_size=1, bias=True, padding_mode='zeros')
        self.act1 = Activation('relu')
        self.conv_expand = nn.Conv2d(in_channels=num_mid, out_channels=num_out, kernel_size=1, bias=True, padding_mode='zeros')
        self.act2 = Activation('hsigmoid')

    def forward(self, x):
        return x * self.act2(self.conv_expand(self.act1(self.conv_reduce(self.pool(x)))))


class ConvUnit(nn.Module):
    def __init__(self, num_in, num_out, kernel_size=1, stride=1, padding=0, num_groups=1, use_act=True, act_type='relu'):
        super().__init__()
        self.conv = nn.Conv2d(in_channels=num_in, out_channels=num_out, kernel_size=kernel_size, stride=stride, padding=padding,
                              groups=num_groups, bias=False, padding_mode='zeros')
        self.bn = nn.BatchNorm2d(num_out)
        self.use_act = use_act
        self.act = Activation(act_type) if use_act else None

    def forward(self, x):
        out = self.bn(self.conv(x))
        return self.act(out) if self.use_act else out


class GhostModule(nn.Module):
    def __init__(self, num_in, num_out, kernel_size=1, stride=1, padding=0, ratio=2, dw_size=3, use_act=True, act_type='relu'):
        super().__init__()
        init_channels = math.ceil(num_out / ratio)
        new_channels = init_channels * (ratio - 1)
        self.primary_conv = ConvUnit(num_in, init_channels, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2,
                                     num_groups=1, use_act=use_act, act_type=act_type)
        self.cheap_operation = ConvUnit(init_channels, new_channels, kernel_size=dw_size, stride=1, padding=dw_size // 2,
                                        num_groups=init_channels, use_act=use_act, act_type=act_type)

    def forward(self, x):
        x1 = self.primary_conv(x)
        return torch.cat([x1, self.cheap_operation(x1)], dim=1)


class GhostModuleMul(nn.Module):
    def __init__(self, num_in, num_out, kernel_size=1, stride=1, padding=0, ratio=2, dw_size=3, use_act=True, act_type='relu'):
        super().__init__()
        self.avgpool2d = nn.AvgPool2d(kernel_size=2, stride=2)
        self.gate_fn = Activation('sigmoid')
        init_channels = math.ceil(num_out / ratio)
        new_channels = init_channels * (ratio - 1)
        self.primary_conv = ConvUnit(num_in, init_channels, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2,
                                     num_groups=1, use_act=use_act, act_type=act_type)
        self.cheap_operation = ConvUnit(init_channels, new_channels, kernel_size=dw_size, stride=1, padding=dw_size // 2,
                                        num_groups=init_channels, use_act=use_act, act_type=act_type)
        self.short_conv = nn.Sequential(
            ConvUnit(num_in, num_out, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2, num_groups=1, use_act=False),
            ConvUnit(num_out, num_out, kernel_size=(1, 5), stride=1, padding=(0, 2), num_groups=num_out, use_act=False),
            ConvUnit(num_out, num_out, kernel_size=(5, 1), stride=1, padding=(2, 0), num_groups=num_out, use_act=False))

    def forward(self, x):
        res = self.gate_fn(self.short_conv(self.avgpool2d(x)))
        x1 = self.primary_conv(x)
        out = torch.cat([x1, self.cheap_operation(x1)], dim=1)
        return out * F.interpolate(res, size=out.shape[-2:], mode='bilinear', align_corners=True)


class Ghostblockv2(_Ref):
    def __init__(self, num_in, num_mid, num_out, kernel_size=3, stride=1, act_type='relu', use_se=False, layer_id=None):
        super().__init__()
        self.use_ori_module = False
        self.ghost1 = GhostModuleMul(num_in, num_mid, kernel_size=1, stride=1, padding=0, act_type=act_type)
        self.use_dw = stride > 1
        self.dw = None
        if self.use_dw:
            self.dw = ConvUnit(num_mid, num_mid, kernel_size=kernel_size, stride=stride, padding=self._get_pad(kernel_size),
                               act_type=act_type, num_groups=num_mid, use_act=False)
        self.use_se = use_se
        if use_se:
            self.se = SE(num_mid)
        self.ghost2 = GhostModule(num_mid, num_out, kernel_size=1, stride=1, padding=0, act_type=act_type, use_act=False)
        self.down_sample = num_in != num_out or stride != 1
        self.shortcut = None
        if self.down_sample:
            self.shortcut = nn.Sequential(
                ConvUnit(num_in, num_in, kernel_size=kernel_size, stride=stride, padding=self._get_pad(kernel_size), num_groups=num_in,
                         use_act=False),
                ConvUnit(num_in, num_out, kernel_size=1, stride=1, padding=0, num_groups=1, use_act=False))

    def _ref(self, x):
        out = self.ghost1(x)
        if self.use_dw:
            out = self.dw(out)
        if self.use_se:
            out = self.se(out)
        out = self.ghost2(out)
        return (self.shortcut(x) if self.down_sample else x) + out

    @staticmethod
    def _get_pad(kernel_size):
        if kernel_size not in (1, 3, 5, 7):
            raise NotImplementedError
        return kernel_size // 2


class C3GhostV2(C3):
    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__(c1, c2, n, shortcut, g, e)
        self.c1_ = 16
        self.c2_ = 16 * e
        c_ = int(c2 * e)
        self.m = nn.Sequential(*(Ghostblockv2(c_, self.c1_, c_) for _ in range(n)))
