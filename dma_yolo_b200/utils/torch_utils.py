"""The few helpers of the reference's utils/torch_utils.py that sit on the detection-forward path."""
from __future__ import annotations

import time

import torch
import torch.nn as nn


def time_sync():
    """utils/torch_utils.py:86-90 — wall clock after a device synchronise."""
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return time.time()


def initialize_weights(model):
    """utils/torch_utils.py:161-170 — BN eps/momentum and in-place activations (conv init untouched)."""
    for m in model.modules():
        t = type(m)
        if t is nn.BatchNorm2d:
            m.eps = 1e-3
            m.momentum = 0.03
        elif t in (nn.Hardswish, nn.LeakyReLU, nn.ReLU, nn.ReLU6, nn.SiLU):
            m.inplace = True


def fuse_conv_and_bn(conv, bn):
    """utils/torch_utils.py:198-218 — W' = diag(g/sqrt(var+eps)) W ; b' = beta - g*mean/sqrt(var+eps) (+ scaled conv bias)."""
    fused = nn.Conv2d(conv.in_channels, conv.out_channels, kernel_size=conv.kernel_size, stride=conv.stride,
                      padding=conv.padding, groups=conv.groups, bias=True).requires_grad_(False).to(conv.weight.device)
    inv = bn.weight.div(torch.sqrt(bn.eps + bn.running_var))
    fused.weight.copy_(torch.mm(torch.diag(inv), conv.weight.clone().view(conv.out_channels, -1)).view(fused.weight.shape))
    b_conv = torch.zeros(conv.weight.size(0), device=conv.weight.device) if conv.bias is None else conv.bias
    b_bn = bn.bias - bn.weight.mul(bn.running_mean).div(torch.sqrt(bn.running_var + bn.eps))
    fused.bias.copy_(torch.mm(torch.diag(inv), b_conv.reshape(-1, 1)).reshape(-1) + b_bn)
    return fused


def model_info(model, verbose=False, img_size=640):
    n_p = sum(x.numel() for x in model.parameters())
    n_g = sum(x.numel() for x in model.parameters() if x.requires_grad)
    return len(list(model.modules())), n_p, n_g


def copy_attr(a, b, include=(), exclude=()):
    for k, v in b.__dict__.items():
        if (len(include) and k not in include) or k.startswith('_') or k in exclude:
            continue
        setattr(a, k, v)


def scale_img(img, ratio=1.0, same_shape=False, gs=32):
    """utils/torch_utils.py scale_img — used only by TTA (_forward_augment)."""
    import math
    import torch.nn.functional as F
    if ratio == 1.0:
        return img
    h, w = img.shape[2:]
    s = (int(h * ratio), int(w * ratio))
    img = F.interpolate(img, size=s, mode='bilinear', align_corners=False)
    if not same_shape:
        h, w = (math.ceil(x * ratio / gs) * gs for x in (h, w))
    return F.pad(img, [0, w - s[1], 0, h - s[0]], value=0.447)
