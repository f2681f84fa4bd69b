"""Drop-in mirror of the hot-path classes of the reference's models/common.py.

Class names, constructor signatures, attribute names and state_dict keys are the reference's
(SURVEY.md 8b / Appendix B), so `parse_model`, YAML configs and pickled checkpoints keep working.
Every class has two bodies:
  * the reference torch-op body — used in training mode or on CPU tensors (the stride probe in
    Model.__init__ is a CPU train-mode forward, models/yolo.py:161-170);
  * `forward_b200` — CUDA eval: bf16 NHWC activations on the sm_100a kernels of libdmayolo.so
    (BN folded into the conv epilogue, residual adds fused, concatenations written in place).
    It raises if the shared library is missing; there is no silent fallback.
"""
from __future__ import annotations

import contextlib
import math
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..ops import ACT_HSWISH, ACT_NONE, ACT_SIGMOID, ACT_SILU, Up


class _State:
    enabled = True  # False -> reference torch bodies everywhere (debug / A-B comparisons only)
    fuse_scconv_gate = True  # SCConv gate in the k3 conv epilogue instead of the stand-alone gate kernel


def set_backend(name: str):
    """'b200' (default): CUDA eval runs on the kernels.  'torch': reference torch-op bodies."""
    assert name in ("b200", "torch")
    _State.enabled = name == "b200"


@contextlib.contextmanager
def reference_ops():
    """Inside this context every module of the package runs its reference torch-op body (used by the torch-only module
    mirrors of models/extra.py so that a whole subtree stays in one dtype / layout)."""
    prev = _State.enabled
    _State.enabled = False
    try:
        yield
    finally:
        _State.enabled = prev


def _first_tensor(x):
    t = x[0] if isinstance(x, (list, tuple)) else x
    return t.src if isinstance(t, Up) else t


def kernel_path(mod: nn.Module, x) -> bool:
    return _State.enabled and not mod.training and _first_tensor(x).is_cuda


def _materialize(x):
    if isinstance(x, Up):
        return x.materialize()
    if isinstance(x, (list, tuple)):
        return [_materialize(t) for t in x]
    return x


def torch_body(mod: nn.Module, fn, x):
    """Run a torch-op body on CUDA for a module that has no kernel path (dtype-matched to its params)."""
    p = next(mod.parameters(), None)
    dt = p.dtype if p is not None else torch.float32
    x = _materialize(x)
    cast = lambda t: t if t.dtype == dt else t.to(dt)
    return fn([cast(t) for t in x] if isinstance(x, list) else cast(x))


def autopad(k, p=None):
    """models/common.py:33-48 — 'same' padding."""
    if p is None:
        p = k // 2 if isinstance(k, int) else [x // 2 for x in k]
    return p


def _act_code(act: nn.Module):
    if isinstance(act, nn.SiLU):
        return ACT_SILU
    if isinstance(act, nn.Identity):
        return ACT_NONE
    if isinstance(act, nn.Hardswish):
        return ACT_HSWISH
    if isinstance(act, nn.Sigmoid):
        return ACT_SIGMOID
    return None


def _ver(*ts):
    return tuple((t.data_ptr(), t._version) for t in ts if t is not None)


def _conv_supported(conv: nn.Conv2d) -> bool:
    return (conv.groups == 1 and tuple(conv.dilation) == (1, 1) and conv.stride[0] == conv.stride[1]
            and conv.padding[0] == conv.padding[1] and isinstance(conv.padding, tuple)
            and conv.padding_mode == 'zeros')


def get_conv_pack(owner: nn.Module, slot: str, conv: nn.Conv2d, bn, device, colscale=None, cols=None, plain=False, spd=False):
    """Cached ConvPack for (conv, bn); rebuilt when any parameter/buffer changed (data_ptr/_version).
    colscale = ((channels, weight), ...): the input-channel ranges of the weight are multiplied by `weight` in fp32 before
    the bf16 rounding (the BiFPN weights of a virtual concat folded into its 1x1 consumer, see ops.VCat).
    cols = ((start, stop), ...): only these input-channel ranges are kept (the K segment of one group of concat parts);
    plain: no BN / bias (scale 1, bias 0) -- the partial sums W0.x0 of ops.VCat.split.
    spd: the 1x1 weight over the 4C channels of a space_to_depth output re-laid as the 2x2 / stride-2 kernel over its C-channel
    input (channel block q = dy + 2 dx, models/common.py:1457-1458; see ops.SPDView)."""
    key = (str(device),) + _ver(conv.weight, conv.bias, *((bn.weight, bn.bias, bn.running_mean, bn.running_var)
                                                          if bn is not None else ())) + ((bn.eps,) if bn is not None else ())
    if colscale is not None or cols is not None or plain or spd:
        key = key + (colscale, cols, plain, spd)
    cache = owner.__dict__.setdefault('_b200_packs', {})
    pk = cache.get(slot)
    if pk is None or pk.key != key:
        wt = conv.weight
        if colscale is not None:
            col = torch.cat([torch.full((c,), w, dtype=torch.float32) for c, w in colscale]).to(wt.device)
            wt = wt.detach().float() * col.view(1, -1, 1, 1)
        if cols is not None:
            wt = torch.cat([wt[:, a:b] for a, b in cols], 1)
        stride, pad = conv.stride[0], conv.padding[0]
        if spd:
            co, c4 = wt.shape[:2]
            wt = wt.detach().reshape(co, 2, 2, c4 // 4).permute(0, 3, 2, 1)   # [co, dx, dy, c] -> [co, c, kh = dy, kw = dx]
            stride, pad = 2, 0
        pk = ops.pack_conv(wt, bn=None if plain else bn, conv_bias=None if plain else conv.bias, stride=stride,
                           pad=pad, device=device, allow_stem_spd=False if spd else True)
        pk.key = key
        cache[slot] = pk
    return pk


class _PackMixin:
    """Packs are derived state: never pickled / deep-copied / saved, dropped on .to()/.half()/load_state_dict."""

    def __getstate__(self):
        d = self.__dict__.copy()
        d.pop('_b200_packs', None)
        return d

    def _apply(self, fn, *a, **k):
        self.__dict__.pop('_b200_packs', None)
        return super()._apply(fn, *a, **k)


class Conv(_PackMixin, nn.Module):
    """Standard convolution conv+BN+act — models/common.py:50-77."""

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p), groups=g, bias=False)
        self.bn = nn.BatchNorm2d(c2)
        self.act = nn.SiLU() if act is True else (act if isinstance(act, nn.Module) else nn.Identity())

    def forward(self, x):
        if kernel_path(self, x):
            return self.forward_b200(x)
        return self.act(self.bn(self.conv(x)))

    def forward_fuse(self, x):
        if kernel_path(self, x):
            return self.forward_b200(x)
        return self.act(self.conv(x))

    def forward_b200(self, x, out=None, residual=None, pool4=False):
        """pool4: the next layer is an SCConv with AvgPool2d(4): ask the kernel for the pooled output as a by-product."""
        code = _act_code(self.act)
        bn = getattr(self, 'bn', None)
        if isinstance(x, ops.VCat):   # a 1x1 layer reads the parts of a concat in place (weights folded into its columns)
            c = self.conv
            srcs = x.sources() if (code is not None and _conv_supported(c) and c.kernel_size == (1, 1) and c.stride == (1, 1)
                                   and c.padding == (0, 0)) else None
            if srcs is not None:
                cs = x.colscale()
                pk = get_conv_pack(self, 'conv@vcat' if cs is not None else 'conv', c, bn, srcs[0].device, cs)
                return ops.conv(srcs, pk, code, out=out, residual=residual)
            x = x.materialize()
        if not _conv_supported(self.conv):
            y = torch_body(self, (lambda t: self.act(bn(self.conv(t)))) if bn is not None else (lambda t: self.act(self.conv(t))), x)
            if residual is not None:
                y = y + residual.to(y.dtype)
            if out is not None:
                out.copy_(y)
                return out
            return y
        pk = get_conv_pack(self, 'conv', self.conv, bn, x.device)
        if code is None:
            y = ops.conv(x, pk, ACT_NONE)
            y = self.act(y)
            if residual is not None:
                y = y + residual
            if out is not None:
                out.copy_(y)
                return out
            return y
        return ops.conv(x, pk, code, out=out, residual=residual, pool4=pool4 and residual is None)


class DWConv(Conv):
    """Depth-wise convolution — models/common.py:79-82 (grouped: stays on torch ops)."""

    def __init__(self, c1, c2, k=1, s=1, act=True):
        super().__init__(c1, c2, k, s, g=math.gcd(c1, c2), act=act)


class Focus(nn.Module):
    """Focus wh information into c-space — models/common.py:84-95."""

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, act=True):
        super().__init__()
        self.conv = Conv(c1 * 4, c2, k, s, p, g, act)

    def forward(self, x):
        if kernel_path(self, x):
            c = x.shape[1]
            if c % 8 == 0 and x.dtype == torch.bfloat16:
                return self.conv.forward_b200(ops.spd(x))
            # an image (3 channels, any input dtype): the prep kernel writes the pixel-unshuffled NHWC bf16 form directly,
            # channel (dy + 2 dx) * C + c as the torch.cat below, zero-padded to the conv pack's channel count
            if c * 4 <= 16 and _conv_supported(self.conv.conv) and self.conv.conv.groups == 1:
                mul = 1.0 / 255.0 if x.dtype == torch.uint8 else 1.0
                return self.conv.forward_b200(ops.input_prep(x, spd=True, cpad=ops.round_up(4 * c, 16), mul=mul))
            return self.conv.forward_b200(ops.spd(ops.as_act(x)))
        return self.conv(torch.cat([x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]], 1))


class Bottleneck(nn.Module):
    """Standard bottleneck — models/common.py:119-137."""

    def __init__(self, c1, c2, shortcut=True, g=1, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_, c2, 3, 1, g=g)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        if kernel_path(self, x):
            return self.forward_b200(x)
        return x + self.cv2(self.cv1(x)) if self.add else self.cv2(self.cv1(x))

    def forward_b200(self, x, out=None):
        x = ops.as_act(x)
        return self.cv2.forward_b200(self.cv1.forward_b200(x), out=out, residual=x if self.add else None)


class BottleneckCSP(nn.Module):
    """CSP Bottleneck — models/common.py:139-157."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = nn.Conv2d(c1, c_, 1, 1, bias=False)
        self.cv3 = nn.Conv2d(c_, c_, 1, 1, bias=False)
        self.cv4 = Conv(2 * c_, c2, 1, 1)
        self.bn = nn.BatchNorm2d(2 * c_)
        self.act = nn.SiLU()
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, g, e=1.0) for _ in range(n)))

    def forward(self, x):
        if kernel_path(self, x):
            return torch_body(self, self._body, x)
        return self._body(x)

    def _body(self, x):
        y1 = self.cv3(self.m(self.cv1(x)))
        y2 = self.cv2(x)
        return self.cv4(self.act(self.bn(torch.cat((y1, y2), dim=1))))


def _run_chain(mods, t, final_out):
    """Run a Sequential of blocks; the last one writes into `final_out` when it knows how to."""
    mods = list(mods)
    for i, b in enumerate(mods):
        last = i == len(mods) - 1
        if hasattr(b, 'forward_b200'):
            t = b.forward_b200(t, out=final_out if last else None) if last else b.forward_b200(t)
        else:
            t = b(t)
            if last and final_out is not None:
                final_out.copy_(ops.as_act(t))
                t = final_out
    return t


class C3(_PackMixin, nn.Module):
    """CSP Bottleneck with 3 convolutions — models/common.py:159-182."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, g, e=1.0) for _ in range(n)))

    def forward(self, x):
        if kernel_path(self, x):
            return self.forward_b200(x)
        return self.cv3(torch.cat((self.m(self.cv1(x)), self.cv2(x)), dim=1))

    def _merged_cv12(self, device, colscale=None, cols=None, plain=False, spd=False):
        """cv1 and cv2 read the same x with the same geometry: ONE GEMM with their weights stacked along Cout
        writes both halves of the concat slab (x is read once, one launch less).  None when they differ.
        colscale: per-part weights of a virtual concat input, folded into the weight columns (get_conv_pack)."""
        a, b = self.cv1, self.cv2
        if not (type(a) is Conv and type(b) is Conv and _conv_supported(a.conv) and _conv_supported(b.conv)):
            return None
        ca, cb = a.conv, b.conv
        code = _act_code(a.act)
        if (code is None or code != _act_code(b.act) or ca.weight.shape != cb.weight.shape or ca.stride != cb.stride
                or ca.padding != cb.padding or ca.out_channels % 16 or (getattr(a, 'bn', None) is None) != (getattr(b, 'bn', None) is None)):
            return None
        sfx = ('' if colscale is None else '@vcat') + ('' if cols is None else '@lo' if plain else '@hi') + ('@spd' if spd else '')
        pa = get_conv_pack(a, 'conv' + sfx, ca, getattr(a, 'bn', None), device, colscale, cols, plain, spd)
        pb = get_conv_pack(b, 'conv' + sfx, cb, getattr(b, 'bn', None), device, colscale, cols, plain, spd)
        cache = self.__dict__.setdefault('_b200_packs', {})
        pk = cache.get('cv12' + sfx)
        if pk is None or pk.key != (pa.key, pb.key):
            pk = ops.ConvPack(w=torch.cat((pa.w, pb.w), 0).contiguous(), scale=torch.cat((pa.scale, pb.scale)).contiguous(),
                              bias=torch.cat((pa.bias, pb.bias)).contiguous(), cin=pa.cin, cin_pad=pa.cin_pad,
                              cout=pa.cout + pb.cout, cout_pad=pa.cout_pad + pb.cout_pad, kh=pa.kh, kw=pa.kw,
                              stride=pa.stride, pad=pa.pad, key=(pa.key, pb.key))
            cache['cv12' + sfx] = pk
        return pk, code

    def _cv12_slab(self, x):
        """[cv1(x) | cv2(x)] as ONE GEMM into a fresh concat slab -> (slab, x) ; (None, x as a tensor) when cv1 / cv2 cannot be
        merged.  x may be lazy: an `ops.SPDView` (space_to_depth -> 1x1 == 2x2 / stride-2 over the un-shuffled tensor: the
        GEMM reads it through the im2col maps) or an `ops.VCat` (a concat that was never written: the K loop walks its parts,
        BiFPN weights folded into the weight columns); both are materialised when the merged GEMM does not apply."""
        c = self.cv1.conv
        one_by_one = c.kernel_size == (1, 1) and c.stride == (1, 1) and c.padding == (0, 0)
        if isinstance(x, ops.SPDView):
            merged = self._merged_cv12(x.src.device, spd=True) if one_by_one and x.src.shape[1] % 16 == 0 else None
            if merged is not None:
                n, _, h, w = x.shape
                slab = ops.empty_nhwc(n, 2 * c.out_channels, h, w, x.src.device)
                ops.conv(x.src, merged[0], merged[1], out=slab)
                return slab, x
            x = x.materialize()
        if isinstance(x, ops.VCat):
            dev = x.src.device
            if one_by_one and all(p.shape[1] % 64 == 0 for p in x.parts) and self._merged_cv12(dev, x.colscale()) is not None:
                n, _, h, w = x.shape
                slab = ops.empty_nhwc(n, 2 * c.out_channels, h, w, dev)
                sp = x.split() if _act_code(self.cv1.act) == ACT_SILU else None
                if sp is not None:
                    # [Up(x0) | x1]: the x0 columns run at x0's own resolution into fp32 partial sums (a quarter of the pixels),
                    # the main GEMM walks the same-resolution parts only and adds up(partial) to its accumulators
                    lo, hi, (lo_cols, hi_cols) = sp
                    p_lo = self._merged_cv12(dev, x.colscale(), lo_cols, True)[0]
                    p_hi = self._merged_cv12(dev, x.colscale(), hi_cols, False)[0]
                    part = ops.conv(lo if len(lo) > 1 else lo[0], p_lo, ACT_NONE, out_fp32=True)
                    ops.conv(hi if len(hi) > 1 else hi[0], p_hi, ACT_SILU, out=slab, pre=part)
                else:
                    merged = self._merged_cv12(dev, x.colscale())
                    ops.conv(x.sources(), merged[0], merged[1], out=slab)
                return slab, x
            x = x.materialize()
        x = ops.as_act(x)
        merged = self._merged_cv12(x.device)
        if merged is None:
            return None, x
        n, _, h, w = x.shape
        slab = ops.empty_nhwc(n, 2 * c.out_channels, h, w, x.device)
        ops.conv(x, merged[0], merged[1], out=slab)          # [cv1(x) | cv2(x)] in one launch
        return slab, x

    def forward_b200(self, x, out=None):
        # cv1 -> bottlenecks write the first half of the concat slab, cv2 the second half: no torch.cat
        c_ = self.cv1.conv.out_channels
        if isinstance(self.m, nn.Sequential):
            slab, x = self._cv12_slab(x)
            if slab is not None:
                if len(self.m) > 0:
                    _run_chain(self.m, slab[:, :c_], slab[:, :c_])   # the last block overwrites cv1(x) in place
                return self.cv3.forward_b200(slab, out=out)
        x = ops.as_act(_materialize(x))
        n, _, h, w = x.shape
        slab = ops.empty_nhwc(n, 2 * c_, h, w, x.device)
        first = slab[:, :c_]
        if isinstance(self.m, nn.Sequential) and len(self.m) > 0:
            _run_chain(self.m, self.cv1.forward_b200(x), first)
        elif isinstance(self.m, nn.Sequential):
            self.cv1.forward_b200(x, out=first)
        else:  # C3 subclasses with a non-Sequential inner module (torch body)
            t = self.m(self.cv1.forward_b200(x))
            first.copy_(ops.as_act(t))
        self.cv2.forward_b200(x, out=slab[:, c_:])
        return self.cv3.forward_b200(slab, out=out)


class SPP(nn.Module):
    """Spatial Pyramid Pooling — models/common.py:212-227."""

    def __init__(self, c1, c2, k=(5, 9, 13)):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * (len(k) + 1), c2, 1, 1)
        self.m = nn.ModuleList([nn.MaxPool2d(kernel_size=x, stride=1, padding=x // 2) for x in k])

    def forward(self, x):
        if kernel_path(self, x):
            return self.forward_b200(x)
        x = self.cv1(x)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            return self.cv2(torch.cat([x] + [m(x) for m in self.m], 1))

    def forward_b200(self, x, out=None):
        x = ops.as_act(x)
        n, _, h, w = x.shape
        c_ = self.cv1.conv.out_channels
        ks = [m.kernel_size if isinstance(m.kernel_size, int) else m.kernel_size[0] for m in self.m]
        slab = ops.empty_nhwc(n, c_ * (len(ks) + 1), h, w, x.device)
        x1 = self.cv1.forward_b200(x, out=slab[:, :c_])
        if len(ks) == 3 and ks[1] == 2 * ks[0] - 1 and ks[2] == 3 * ks[0] - 2:  # (5,9,13): one cascaded pass
            ops.sppf_pool3(x1, slab[:, c_:2 * c_], slab[:, 2 * c_:3 * c_], slab[:, 3 * c_:], ks[0])
        else:
            for i, k in enumerate(ks):
                ops.maxpool_s1(x1, k, out=slab[:, (i + 1) * c_:(i + 2) * c_])
        return self.cv2.forward_b200(slab, out=out)


class SPPF(nn.Module):
    """Spatial Pyramid Pooling - Fast — models/common.py:243-258."""

    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def forward(self, x):
        if kernel_path(self, x):
            return self.forward_b200(x)
        x = self.cv1(x)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            y1 = self.m(x)
            y2 = self.m(y1)
            return self.cv2(torch.cat([x, y1, y2, self.m(y2)], 1))

    def forward_b200(self, x, out=None):
        x = ops.as_act(x)
        n, _, h, w = x.shape
        c_ = self.cv1.conv.out_channels
        k = self.m.kernel_size if isinstance(self.m.kernel_size, int) else self.m.kernel_size[0]
        slab = ops.empty_nhwc(n, 4 * c_, h, w, x.device)
        x1 = self.cv1.forward_b200(x, out=slab[:, :c_])
        ops.sppf_pool3(x1, slab[:, c_:2 * c_], slab[:, 2 * c_:3 * c_], slab[:, 3 * c_:], k)
        return self.cv2.forward_b200(slab, out=out)


class Contract(nn.Module):
    """Contract width-height into channels — models/common.py:357-369."""

    def __init__(self, gain=2):
        super().__init__()
        self.gain = gain

    def forward(self, x):
        x = _materialize(x)
        b, c, h, w = x.size()
        s = self.gain
        x = x.view(b, c, h // s, s, w // s, s) if x.is_contiguous() else x.reshape(b, c, h // s, s, w // s, s)
        x = x.permute(0, 3, 5, 1, 2, 4).contiguous()
        return x.view(b, c * s * s, h // s, w // s)


class Expand(nn.Module):
    """Expand channels into width-height — models/common.py:371-383."""

    def __init__(self, gain=2):
        super().__init__()
        self.gain = gain

    def forward(self, x):
        x = _materialize(x)
        b, c, h, w = x.size()
        s = self.gain
        x = x.reshape(b, s, s, c // s ** 2, h, w)
        x = x.permute(0, 3, 4, 1, 5, 2).contiguous()
        return x.view(b, c // s ** 2, h * s, w * s)


class Concat(nn.Module):
    """Concatenate a list of tensors along dimension — models/common.py:656-664."""

    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def forward(self, x):
        if kernel_path(self, x) and self.d == 1 and all(_is_act_like(t) for t in x):
            return ops.vcat(x, (1.0,) * len(x)) if 2 <= len(x) <= 3 else ops.concat(x)
        return torch.cat(_materialize(x), self.d)


def _is_act_like(t):
    t = t.src if isinstance(t, Up) else t
    return t.dim() == 4 and t.shape[1] % 8 == 0


class _AdWeights:
    """Normalised BiFPN weights w/(sum w + eps), cached on the host per parameter version (one D2H)."""

    def _norm_weights(self):
        key = _ver(self.w)
        c = self.__dict__.get('_b200_w')
        if c is None or c[0] != key:
            w = self.w.detach().float()
            c = (key, (w / (torch.sum(w, dim=0) + self.epsilon)).tolist())
            self.__dict__['_b200_w'] = c
        return c[1]

    def __getstate__(self):
        d = self.__dict__.copy()
        d.pop('_b200_w', None)
        return d


class AdConcat2(_AdWeights, nn.Module):
    """BiFPN learned-weight fusion of two branches, concatenated — models/common.py:994-1008."""

    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension
        self.w = nn.Parameter(torch.ones(2, dtype=torch.float32), requires_grad=True)
        self.epsilon = 0.0001

    def forward(self, x):
        if kernel_path(self, x) and self.d == 1:
            return ops.vcat(x, self._norm_weights())
        x = _materialize(x)
        w = self.w
        weight = w / (torch.sum(w, dim=0) + self.epsilon)
        x = [weight[0] * x[0], weight[1] * x[1]]
        return torch.cat(x, self.d)


class AdConcat3(_AdWeights, nn.Module):
    """Three-branch variant — models/common.py:1010-1026."""

    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension
        self.w = nn.Parameter(torch.ones(3, dtype=torch.float32), requires_grad=True)
        self.epsilon = 0.0001

    def forward(self, x):
        if kernel_path(self, x) and self.d == 1:
            return ops.vcat(x, self._norm_weights())
        x = _materialize(x)
        w = self.w
        weight = w / (torch.sum(w, dim=0) + self.epsilon)
        x = [weight[0] * x[0], weight[1] * x[1], weight[2] * x[2]]
        return torch.cat(x, self.d)


class Adapt_Add2(_AdWeights, nn.Module):
    """Learned-weight add of two branches + SiLU — models/common.py:1028-1045."""

    def __init__(self):
        super().__init__()
        self.w = nn.Parameter(torch.ones(2, dtype=torch.float32), requires_grad=True)
        self.epsilon = 0.0001
        self.silu = nn.SiLU()

    def forward(self, x):
        x = _materialize(x)
        if kernel_path(self, x):
            return ops.adapt_add(x, self._norm_weights())
        w = self.w
        weight = w / (torch.sum(w, dim=0) + self.epsilon)
        return self.silu(weight[0] * x[0] + weight[1] * x[1])


class Adapt_Add3(_PackMixin, _AdWeights, nn.Module):
    """Three-branch add with a shared 1x1 conv on the first two — models/common.py:1047-1061."""

    def __init__(self, d1, d2, d3):
        super().__init__()
        self.w = nn.Parameter(torch.ones(3, dtype=torch.float32), requires_grad=True)
        self.epsilon = 0.0001
        self.conv = nn.Conv2d(d1, d3, kernel_size=1, stride=1, padding=0)
        self.silu = nn.SiLU()

    def forward(self, x):
        x = _materialize(x)
        if kernel_path(self, x):
            pk = get_conv_pack(self, 'conv', self.conv, None, x[0].device)
            a, b = ops.conv(x[0], pk, ACT_NONE), ops.conv(x[1], pk, ACT_NONE)
            return ops.adapt_add([a, b, x[2]], self._norm_weights())
        w = self.w
        weight = w / (torch.sum(w, dim=0) + self.epsilon)
        return self.silu(weight[0] * self.conv(x[0]) + weight[1] * self.conv(x[1]) + weight[2] * x[2])


class CoorAttention(nn.Module):
    """Coordinate Attention — models/common.py:1158-1207."""

    def __init__(self, c1, c2, reduction=32):
        super().__init__()
        self.pool_h = nn.AdaptiveAvgPool2d((None, 1))
        self.pool_w = nn.AdaptiveAvgPool2d((1, None))
        c_ = max(8, c1 // reduction)
        self.conv1 = nn.Conv2d(c1, c_, kernel_size=1, stride=1, padding=0)
        self.bn1 = nn.BatchNorm2d(c_)
        self.act = nn.Hardswish()
        self.conv_w = nn.Conv2d(c_, c2, kernel_size=1, stride=1, padding=0)
        self.conv_h = nn.Conv2d(c_, c2, kernel_size=1, stride=1, padding=0)

    def forward(self, x):
        if kernel_path(self, x):
            return self.forward_b200(x)
        identity = x
        n, c, h, w = x.size()
        x_h = self.pool_h(x)
        x_w = self.pool_w(x).permute(0, 1, 3, 2)
        y = torch.cat([x_h, x_w], dim=2)
        y = self.act(self.bn1(self.conv1(y)))
        x_h, x_w = torch.split(y, [h, w], dim=2)
        x_w = x_w.permute(0, 1, 3, 2)
        a_h = self.conv_h(x_h).sigmoid()
        a_w = self.conv_w(x_w).sigmoid()
        return identity * a_w * a_h

    def __getstate__(self):
        d = self.__dict__.copy()
        d.pop('_b200_ca', None)
        return d

    def _apply(self, fn, *a, **k):
        self.__dict__.pop('_b200_ca', None)
        return super()._apply(fn, *a, **k)

    def forward_b200(self, x, out=None):
        if not isinstance(self.act, nn.Hardswish) or self.conv_h.out_channels != self.conv1.in_channels:
            raise ops.DmayError("CoorAttention kernel path needs Hardswish and c2 == c1")
        key = (str(x.device),) + _ver(self.conv1.weight, self.conv1.bias, self.bn1.weight, self.bn1.bias,
                                      self.bn1.running_mean, self.bn1.running_var, self.conv_h.weight,
                                      self.conv_h.bias, self.conv_w.weight, self.conv_w.bias) + (self.bn1.eps,)
        pk = self.__dict__.get('_b200_ca')
        if pk is None or pk.key != key:
            pk = ops.pack_coordatt(self.conv1, self.bn1, self.conv_h, self.conv_w, x.device)
            pk.key = key
            self.__dict__['_b200_ca'] = pk
        return ops.coordatt(x, pk, out=out)


CA = CoorAttention  # the YAMLs name `CA`; the reference never defines it


class CABottleneck(nn.Module):
    """Bottleneck + CoordAtt — models/common.py:1209-1227."""

    def __init__(self, c1, c2, shortcut=True, g=1, e=0.5, reduction=32):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_, c2, 3, 1, g=g)
        self.ca = CoorAttention(c2, c2, reduction)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        if kernel_path(self, x):
            return self.forward_b200(x)
        return x + self.ca(self.cv2(self.cv1(x))) if self.add else self.ca(self.cv2(self.cv1(x)))

    def forward_b200(self, x, out=None):
        x = ops.as_act(x)
        y = self.ca.forward_b200(self.cv2.forward_b200(self.cv1.forward_b200(x)), out=None if self.add else out)
        if self.add:  # x + ca(...): learned-weight add kernel is silu-fused, so use the exact 2-input sum
            y = _residual_add(x, y, out)
        return y


def _residual_add(a, b, out=None):
    y = torch.add(a, b) if out is None else torch.add(a, b, out=out)
    return y


class C3CA(C3):
    """C3 with CABottleneck — models/common.py:1229-1235."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__(c1, c2, n, shortcut, g, e)
        c_ = int(c2 * e)
        self.m = nn.Sequential(*(CABottleneck(c_, c_, shortcut, g, e=1.0) for _ in range(n)))


# ---- 8f-1: Swin transformer block inside C3 (C3STR) ---------------------------------------------------------------
class Mlp(nn.Module):
    """MLP as used in Vision Transformer — models/common.py:97-117."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.drop1(self.act(self.fc1(x)))))


class DropPath(nn.Module):
    """Stochastic depth per sample — models/common.py:386-413 (identity in eval)."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if not self.drop_prob or not self.training:
            return x
        keep = 1 - self.drop_prob
        rnd = keep + torch.rand((x.shape[0],) + (1,) * (x.ndim - 1), dtype=x.dtype, device=x.device)
        return x.div(keep) * rnd.floor_()


def window_partition(x, window_size: int):
    """(B, H, W, C) -> (num_windows*B, ws, ws, C) — models/common.py:415-428."""
    B, H, W, C = x.shape
    x = x.view(B, H // window_size, window_size, W // window_size, window_size, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, window_size, window_size, C)


def window_reverse(windows, window_size: int, H: int, W: int):
    """models/common.py:430-446."""
    B = int(windows.shape[0] / (H * W / window_size / window_size))
    x = windows.view(B, H // window_size, W // window_size, window_size, window_size, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


class WindowAttention(nn.Module):
    """Window based multi-head self attention with relative position bias — models/common.py:448-515."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.dim = dim
        self.window_size = window_size
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * window_size[0] - 1) * (2 * window_size[1] - 1), num_heads))
        coords = torch.stack(torch.meshgrid([torch.arange(window_size[0]), torch.arange(window_size[1])], indexing='ij'))
        cf = torch.flatten(coords, 1)
        rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += window_size[0] - 1
        rel[:, :, 1] += window_size[1] - 1
        rel[:, :, 0] *= 2 * window_size[1] - 1
        self.register_buffer('relative_position_index', rel.sum(-1))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        self.softmax = nn.Softmax(dim=-1)

    def rel_bias(self):
        """[heads, N, N] relative position bias (gathered table)."""
        n = self.window_size[0] * self.window_size[1]
        return self.relative_position_bias_table[self.relative_position_index.view(-1)].view(n, n, -1).permute(2, 0, 1).contiguous()

    def forward(self, x, mask=None):
        B_, N, C = x.shape
        qkv = self.qkv(x).reshape(B_, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = (q * self.scale) @ k.transpose(-2, -1)
        attn = attn + self.rel_bias().unsqueeze(0)
        if mask is not None:
            nW = mask.shape[0]
            attn = attn.view(B_ // nW, nW, self.num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
            attn = attn.view(-1, self.num_heads, N, N)
        attn = self.attn_drop(self.softmax(attn))
        x = (attn @ v).transpose(1, 2).reshape(B_, N, C)
        return self.proj_drop(self.proj(x))


def _linear_pack(owner, slot, lin: nn.Linear, device):
    """A Linear as a 1x1 convolution for the implicit-GEMM kernel (cached like the conv packs)."""
    key = (str(device),) + _ver(lin.weight, lin.bias)
    cache = owner.__dict__.setdefault('_b200_packs', {})
    pk = cache.get(slot)
    if pk is None or pk.key != key:
        pk = ops.pack_conv(lin.weight.view(lin.out_features, lin.in_features, 1, 1), conv_bias=lin.bias, device=device)
        pk.key = key
        cache[slot] = pk
    return pk


class SwinTransformerLayer(_PackMixin, nn.Module):
    """models/common.py:517-634.  Note the reference reads its NCHW input as (b, c, w, h): windows, shift mask and
    relative positions live in the TRANSPOSED spatial frame; the kernel path reproduces that by index arithmetic."""

    def __init__(self, c, num_heads, window_size=7, shift_size=0, mlp_ratio=4, qkv_bias=False, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        if num_heads > 10:
            drop_path = 0.1
        self.window_size = window_size
        self.shift_size = shift_size
        self.mlp_ratio = mlp_ratio
        self.norm1 = norm_layer(c)
        self.attn = WindowAttention(c, window_size=(window_size, window_size), num_heads=num_heads, qkv_bias=qkv_bias,
                                    attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(c)
        self.mlp = Mlp(in_features=c, hidden_features=int(c * mlp_ratio), act_layer=act_layer, drop=drop)

    def create_mask(self, x, H, W):
        """models/common.py:567-591, including its quirk: the first h-slice is the TUPLE (0, -ws), i.e. the two rows
        0 and Hp-ws, not a range."""
        ws, ss = self.window_size, self.shift_size
        Hp = int(math.ceil(H / ws)) * ws
        Wp = int(math.ceil(W / ws)) * ws
        img_mask = torch.zeros((1, Hp, Wp, 1), device=x.device)
        h_slices = ((0, -ws), slice(-ws, -ss), slice(-ss, None))
        w_slices = (slice(0, -ws), slice(-ws, -ss), slice(-ss, None))
        cnt = 0
        for h in h_slices:
            for w in w_slices:
                img_mask[:, h, w, :] = cnt
                cnt += 1
        mw = window_partition(img_mask, ws).view(-1, ws * ws)
        am = mw.unsqueeze(1) - mw.unsqueeze(2)
        return am.masked_fill(am != 0, -100.0).masked_fill(am == 0, 0.0)

    def forward(self, x):
        if kernel_path(self, x) and self._kernel_ok():
            return self.forward_b200(x)
        b, c, w, h = x.shape
        x = x.permute(0, 3, 2, 1).contiguous()
        attn_mask = self.create_mask(x, h, w)
        shortcut = x
        x = self.norm1(x)
        ws = self.window_size
        pad_r = (ws - w % ws) % ws
        pad_b = (ws - h % ws) % ws
        x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b))
        _, hp, wp, _ = x.shape
        if self.shift_size > 0:
            shifted = torch.roll(x, shifts=(-self.shift_size, -self.shift_size), dims=(1, 2))
        else:
            shifted, attn_mask = x, None
        xw = window_partition(shifted, ws).view(-1, ws * ws, c)
        aw = self.attn(xw, mask=attn_mask).view(-1, ws, ws, c)
        shifted = window_reverse(aw, ws, hp, wp)
        x = torch.roll(shifted, shifts=(self.shift_size, self.shift_size), dims=(1, 2)) if self.shift_size > 0 else shifted
        if pad_r > 0 or pad_b > 0:
            x = x[:, :h, :w, :].contiguous()
        x = shortcut + self.drop_path(x)
        x = x + self.drop_path(self.mlp(self.norm2(x)))
        return x.permute(0, 3, 2, 1).contiguous()

    def _kernel_ok(self):
        a = self.attn
        return (self.window_size == 8 and a.dim == a.num_heads * 32 and a.qkv.bias is None
                and isinstance(self.norm1, nn.LayerNorm) and isinstance(self.mlp.act, nn.GELU)
                and getattr(self.mlp.act, 'approximate', 'none') == 'none')

    def _aux(self, dev, H, W):
        """fp32 device operands derived from the parameters: LN affine, gathered relative bias, shift mask per (H, W)."""
        a = self.attn
        key = (str(dev),) + _ver(self.norm1.weight, self.norm1.bias, self.norm2.weight, self.norm2.bias,
                                 a.relative_position_bias_table)
        cache = self.__dict__.setdefault('_b200_packs', {})
        aux = cache.get('aux')
        if aux is None or aux['key'] != key:
            f = lambda t: t.detach().float().to(dev).contiguous()
            aux = dict(key=key, g1=f(self.norm1.weight), b1=f(self.norm1.bias), g2=f(self.norm2.weight), b2=f(self.norm2.bias),
                       rel=f(a.rel_bias()), masks={})
            cache['aux'] = aux
        mask = None
        if self.shift_size > 0:
            mask = aux['masks'].get((H, W))
            if mask is None:
                # the reference calls create_mask(x, h, w) with h = x.shape[3], w = x.shape[2] of the NCHW input
                mask = aux['masks'][(H, W)] = self.create_mask(torch.zeros(1), W, H).float().to(dev).contiguous()
        return aux, mask

    def forward_b200(self, x, out=None):
        x = ops.as_act(x)
        n, c, H, W = x.shape
        dev = x.device
        aux, mask = self._aux(dev, H, W)
        a = self.attn
        y = ops.layernorm(x, aux['g1'], aux['b1'], self.norm1.eps)
        qkv = ops.conv(y, _linear_pack(self, 'qkv', a.qkv, dev), ACT_NONE)
        att = ops.window_attention(qkv, aux['rel'], mask, a.num_heads, self.shift_size, a.scale)
        x1 = ops.conv(att, _linear_pack(self, 'proj', a.proj, dev), ACT_NONE, residual=x)
        y2 = ops.layernorm(x1, aux['g2'], aux['b2'], self.norm2.eps)
        hdn = ops.conv(y2, _linear_pack(self, 'fc1', self.mlp.fc1, dev), ops.ACT_GELU)
        return ops.conv(hdn, _linear_pack(self, 'fc2', self.mlp.fc2, dev), ACT_NONE, residual=x1, out=out)


class SwinTransformerBlock(nn.Module):
    """models/common.py:636-654."""

    def __init__(self, c1, c2, num_heads, num_layers, window_size=8):
        super().__init__()
        self.conv = None
        if c1 != c2:
            self.conv = Conv(c1, c2)
        self.window_size = window_size
        self.shift_size = window_size // 2
        self.tr = nn.Sequential(*(SwinTransformerLayer(c2, num_heads=num_heads, window_size=window_size,
                                                       shift_size=0 if (i % 2 == 0) else self.shift_size)
                                  for i in range(num_layers)))

    def forward(self, x):
        if self.conv is not None:
            x = self.conv(x)
        return self.tr(x)

    def forward_b200(self, x, out=None):
        if self.conv is not None:
            x = self.conv.forward_b200(x)
        return _run_chain(self.tr, x, out)


class C3STR(C3):
    """C3 module with SwinTransformerBlock() — models/common.py:191-196."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__(c1, c2, n, shortcut, g, e)
        c_ = int(c2 * e)
        self.m = SwinTransformerBlock(c_, c_, c_ // 32, n)

    def forward_b200(self, x, out=None):
        c_ = self.cv1.conv.out_channels
        slab, x = self._cv12_slab(x)                  # x may be a lazy concat / space_to_depth (see C3._cv12_slab)
        if slab is not None:
            t = slab[:, :c_]
        else:
            n, _, h, w = x.shape
            slab = ops.empty_nhwc(n, 2 * c_, h, w, x.device)
            t = self.cv1.forward_b200(x)
            self.cv2.forward_b200(x, out=slab[:, c_:])
        self.m.forward_b200(t, out=slab[:, :c_])      # the last Swin layer writes the first half of the concat slab
        return self.cv3.forward_b200(slab, out=out)


# ---- 8f-2: HorNet block inside C3 (C3HB) ------------------------------------------------------------------------------
class LayerNorm(nn.Module):
    """LayerNorm with channels_last (default) or channels_first data format — models/common.py:1385-1409."""

    def __init__(self, normalized_shape, eps=1e-6, data_format='channels_last'):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps = eps
        self.data_format = data_format
        if self.data_format not in ['channels_last', 'channels_first']:
            raise NotImplementedError
        self.normalized_shape = (normalized_shape,)

    def forward(self, x):
        if self.data_format == 'channels_last':
            return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        u = x.mean(1, keepdim=True)
        s = (x - u).pow(2).mean(1, keepdim=True)
        x = (x - u) / torch.sqrt(s + self.eps)
        return self.weight[:, None, None] * x + self.bias[:, None, None]


def get_dwconv(dim, kernel, bias):
    """models/common.py:1348-1349."""
    return nn.Conv2d(dim, dim, kernel_size=kernel, padding=(kernel - 1) // 2, bias=bias, groups=dim)


class GnConv(_PackMixin, nn.Module):
    """Recursive gated convolution — models/common.py:1318-1346."""

    def __init__(self, c1, c2, ksize=1, stride=1, order=5, gflayer=None, h=14, w=8, s=1.0):
        super().__init__()
        self.order = order
        self.dims = [c1 // 2 ** i for i in range(order)]
        self.dims.reverse()
        self.proj_in = nn.Conv2d(c1, 2 * c1, 1)
        if gflayer is None:
            self.dwconv = get_dwconv(sum(self.dims), 7, True)
        else:
            self.dwconv = gflayer(sum(self.dims), h=h, w=w)
        self.proj_out = Conv(c1, c2, ksize, stride)
        self.pws = nn.ModuleList([nn.Conv2d(self.dims[i], self.dims[i + 1], 1) for i in range(order - 1)])
        self.scale = s

    def forward(self, x, mask=None, dummy=False):
        if kernel_path(self, x) and self._kernel_ok():
            return self.forward_b200(x)
        fused_x = self.proj_in(x)
        pwa, abc = torch.split(fused_x, (self.dims[0], sum(self.dims)), dim=1)
        dw_list = torch.split(self.dwconv(abc) * self.scale, self.dims, dim=1)
        x = pwa * dw_list[0]
        for i in range(self.order - 1):
            x = self.pws[i](x) * dw_list[i + 1]
        return self.proj_out(x)

    def _kernel_ok(self):
        dw = self.dwconv
        return (isinstance(dw, nn.Conv2d) and dw.kernel_size == (7, 7) and dw.groups == dw.in_channels and dw.padding == (3, 3)
                and dw.stride == (1, 1) and self.dims[0] % 2 == 0 and all(d % 8 == 0 for d in self.dims[1:])
                and self.proj_in.in_channels % 16 == 0)

    def _conv1x1_pack(self, slot, conv, dev):
        key = (str(dev),) + _ver(conv.weight, conv.bias)
        cache = self.__dict__.setdefault('_b200_packs', {})
        pk = cache.get(slot)
        if pk is None or pk.key != key:
            pk = ops.pack_conv(conv.weight, conv_bias=conv.bias, device=dev)
            pk.key = key
            cache[slot] = pk
        return pk

    def forward_b200(self, x, out=None):
        x = ops.as_act(x)
        dev = x.device
        dims = self.dims
        fused = ops.conv(x, self._conv1x1_pack('proj_in', self.proj_in, dev), ACT_NONE)          # [pwa | abc]
        # depth-wise 7x7 over abc; segment i (gating order i) lands at an 8-channel-aligned offset
        seg_start = [sum(dims[:i]) for i in range(len(dims))]
        seg_out, o = [], 0
        for d in dims:
            seg_out.append(o)
            o += ops.round_up(d, 8)
        cache = self.__dict__.setdefault('_b200_packs', {})
        kdw = (str(dev),) + _ver(self.dwconv.weight, self.dwconv.bias)
        dwp = cache.get('dw')
        if dwp is None or dwp[0] != kdw:
            w49 = self.dwconv.weight.detach().float().reshape(sum(dims), 49).t().contiguous().to(dev)
            dwp = cache['dw'] = (kdw, w49, self.dwconv.bias.detach().float().to(dev).contiguous())
        dw = ops.dwconv7(fused, dims[0], dwp[1], dwp[2], self.scale, seg_start, seg_out, o)
        # recursive gating: x0 = pwa * dw0, x_{i+1} = pws_i(x_i) * dw_{i+1} (the product is the GEMM's epilogue)
        t = ops.mul_channels(fused[:, :dims[0]], dw[:, :dims[0]], dims[0], ops.round_up(dims[0], 16))
        n, _, h, w = x.shape
        for i in range(self.order - 1):
            d1 = dims[i + 1]
            buf = torch.zeros((n, h, w, ops.round_up(d1, 16)), device=dev, dtype=torch.bfloat16).permute(0, 3, 1, 2) \
                if d1 % 16 else ops.empty_nhwc(n, d1, h, w, dev)
            ops.conv(t, self._conv1x1_pack(f'pws{i}', self.pws[i], dev), ACT_NONE, out=buf[:, :d1],
                     residual=dw[:, seg_out[i + 1]:seg_out[i + 1] + d1], res_mul=True)
            t = buf
        return self.proj_out.forward_b200(t[:, :dims[-1]] if t.shape[1] != dims[-1] else t, out=out)


class HorBlock(_PackMixin, nn.Module):
    """HorNet block — models/common.py:1351-1383."""

    def __init__(self, dim, drop_path=0., layer_scale_init_value=1e-6, gnconv=GnConv):
        super().__init__()
        self.norm1 = LayerNorm(dim, eps=1e-6, data_format='channels_first')
        self.gnconv = GnConv(dim, dim)
        self.norm2 = LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.gamma1 = nn.Parameter(layer_scale_init_value * torch.ones(dim), requires_grad=True) \
            if layer_scale_init_value > 0 else None
        self.gamma2 = nn.Parameter(layer_scale_init_value * torch.ones(dim), requires_grad=True) \
            if layer_scale_init_value > 0 else None
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()

    def forward(self, x):
        if kernel_path(self, x) and self.gnconv._kernel_ok() and x.shape[1] % 8 == 0:
            return self.forward_b200(x)
        B, C, H, W = x.shape
        gamma1 = self.gamma1.view(C, 1, 1) if self.gamma1 is not None else 1
        x = x + self.drop_path(gamma1 * self.gnconv(self.norm1(x)))
        inp = x
        x = self.pwconv2(self.act(self.pwconv1(self.norm2(x.permute(0, 2, 3, 1)))))
        if self.gamma2 is not None:
            x = self.gamma2 * x
        return inp + self.drop_path(x.permute(0, 3, 1, 2))

    def forward_b200(self, x, out=None):
        x = ops.as_act(x)
        n, c, h, w = x.shape
        dev = x.device
        key = (str(dev),) + _ver(self.norm1.weight, self.norm1.bias, self.norm2.weight, self.norm2.bias, self.gamma1, self.gamma2,
                                 self.pwconv2.weight, self.pwconv2.bias)
        cache = self.__dict__.setdefault('_b200_packs', {})
        aux = cache.get('aux')
        if aux is None or aux['key'] != key:
            f = lambda t: t.detach().float().to(dev).contiguous()
            g2 = (self.gamma2.detach().float() if self.gamma2 is not None else torch.ones(c)).to(dev)
            fc2 = ops.pack_conv(self.pwconv2.weight.view(c, 4 * c, 1, 1), conv_bias=self.pwconv2.bias, device=dev)
            fc2.scale = (fc2.scale[:c] * g2).contiguous()          # gamma2 * (W h + b): layer scale folded into the GEMM's
            fc2.bias = (fc2.bias[:c] * g2).contiguous()            # per-channel scale / bias (c % 16 == 0: no padding rows)
            aux = dict(key=key, g1=f(self.norm1.weight), b1=f(self.norm1.bias), g2=f(self.norm2.weight), b2=f(self.norm2.bias),
                       gamma1=f(self.gamma1) if self.gamma1 is not None else torch.ones(c, device=dev), fc2=fc2)
            cache['aux'] = aux
        y = ops.layernorm(x, aux['g1'], aux['b1'], self.norm1.eps)
        g = self.gnconv.forward_b200(y)
        x1 = ops.axpy_channels(x, g, aux['gamma1'])
        y2 = ops.layernorm(x1, aux['g2'], aux['b2'], self.norm2.eps)
        hdn = ops.conv(y2, _linear_pack(self, 'fc1', self.pwconv1, dev), ops.ACT_GELU)
        return ops.conv(hdn, aux['fc2'], ACT_NONE, residual=x1, out=out)


class C3HB(_PackMixin, nn.Module):
    """CSP bottleneck with HorBlocks — models/common.py:1412-1426."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(HorBlock(c_) for _ in range(n)))

    def forward(self, x):
        if kernel_path(self, x):
            return self.forward_b200(x)
        return self.cv3(torch.cat((self.m(self.cv1(x)), self.cv2(x)), dim=1))

    _merged_cv12 = C3._merged_cv12
    _cv12_slab = C3._cv12_slab

    def forward_b200(self, x, out=None):
        c_ = self.cv1.conv.out_channels
        slab, x = self._cv12_slab(x)                  # x may be a lazy concat / space_to_depth (see C3._cv12_slab)
        if slab is not None:
            t = slab[:, :c_]
        else:
            n, _, h, w = x.shape
            slab = ops.empty_nhwc(n, 2 * c_, h, w, x.device)
            t = self.cv1.forward_b200(x)
            self.cv2.forward_b200(x, out=slab[:, c_:])
        _run_chain(self.m, t, slab[:, :c_])
        return self.cv3.forward_b200(slab, out=out)


class SPPCSPC(nn.Module):
    """CSP SPP — models/common.py:1237-1255."""

    def __init__(self, c1, c2, n=1, shortcut=False, g=1, e=0.5, k=(5, 9, 13)):
        super().__init__()
        c_ = int(2 * c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(c_, c_, 3, 1)
        self.cv4 = Conv(c_, c_, 1, 1)
        self.m = nn.ModuleList([nn.MaxPool2d(kernel_size=x, stride=1, padding=x // 2) for x in k])
        self.cv5 = Conv(4 * c_, c_, 1, 1)
        self.cv6 = Conv(c_, c_, 3, 1)
        self.cv7 = Conv(2 * c_, c2, 1, 1)

    def forward(self, x):
        if kernel_path(self, x):
            return self.forward_b200(x)
        x1 = self.cv4(self.cv3(self.cv1(x)))
        y1 = self.cv6(self.cv5(torch.cat([x1] + [m(x1) for m in self.m], 1)))
        y2 = self.cv2(x)
        return self.cv7(torch.cat((y1, y2), dim=1))

    def forward_b200(self, x, out=None):
        x = ops.as_act(x)
        n, _, h, w = x.shape
        c_ = self.cv1.conv.out_channels
        ks = [m.kernel_size if isinstance(m.kernel_size, int) else m.kernel_size[0] for m in self.m]
        slab4 = ops.empty_nhwc(n, (len(ks) + 1) * c_, h, w, x.device)
        x1 = self.cv4.forward_b200(self.cv3.forward_b200(self.cv1.forward_b200(x)), out=slab4[:, :c_])
        if len(ks) == 3 and ks[1] == 2 * ks[0] - 1 and ks[2] == 3 * ks[0] - 2:
            ops.sppf_pool3(x1, slab4[:, c_:2 * c_], slab4[:, 2 * c_:3 * c_], slab4[:, 3 * c_:], ks[0])
        else:
            for i, k in enumerate(ks):
                ops.maxpool_s1(x1, k, out=slab4[:, (i + 1) * c_:(i + 2) * c_])
        slab2 = ops.empty_nhwc(n, 2 * c_, h, w, x.device)
        self.cv6.forward_b200(self.cv5.forward_b200(slab4), out=slab2[:, :c_])
        self.cv2.forward_b200(x, out=slab2[:, c_:])
        return self.cv7.forward_b200(slab2, out=out)


class SPPFCSPC(nn.Module):
    """CSP SPPF: cascaded 5x5 max-pools — models/common.py:1257-1276."""

    def __init__(self, c1, c2, n=1, shortcut=False, g=1, e=0.5, k=5):
        super().__init__()
        c_ = int(2 * c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(c_, c_, 3, 1)
        self.cv4 = Conv(c_, c_, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)
        self.cv5 = Conv(4 * c_, c_, 1, 1)
        self.cv6 = Conv(c_, c_, 3, 1)
        self.cv7 = Conv(2 * c_, c2, 1, 1)

    def forward(self, x):
        if kernel_path(self, x):
            return self.forward_b200(x)
        x1 = self.cv4(self.cv3(self.cv1(x)))
        x2 = self.m(x1)
        x3 = self.m(x2)
        y1 = self.cv6(self.cv5(torch.cat((x1, x2, x3, self.m(x3)), 1)))
        y2 = self.cv2(x)
        return self.cv7(torch.cat((y1, y2), dim=1))

    def forward_b200(self, x, out=None):
        # cv4 writes x1 into channels [0,c_) of the 4c_ slab cv5 reads; the pool cascade fills the
        # other three quarters in one pass; cv6 / cv2 write the two halves of cv7's input.
        x = ops.as_act(x)
        n, _, h, w = x.shape
        c_ = self.cv1.conv.out_channels
        k = self.m.kernel_size if isinstance(self.m.kernel_size, int) else self.m.kernel_size[0]
        slab4 = ops.empty_nhwc(n, 4 * c_, h, w, x.device)
        x1 = self.cv4.forward_b200(self.cv3.forward_b200(self.cv1.forward_b200(x)), out=slab4[:, :c_])
        ops.sppf_pool3(x1, slab4[:, c_:2 * c_], slab4[:, 2 * c_:3 * c_], slab4[:, 3 * c_:], k)
        slab2 = ops.empty_nhwc(n, 2 * c_, h, w, x.device)
        self.cv6.forward_b200(self.cv5.forward_b200(slab4), out=slab2[:, :c_])
        self.cv2.forward_b200(x, out=slab2[:, c_:])
        return self.cv7.forward_b200(slab2, out=out)


class SCConv(_PackMixin, nn.Module):
    """Self-calibrated convolution (SCNet, CVPR'20) — models/common.py:1279-1316."""

    def __init__(self, c1, c2, stride, groups=1, dilation=1, pooling_r=4):
        super().__init__()
        self.k2 = nn.Sequential(
            nn.AvgPool2d(kernel_size=pooling_r, stride=pooling_r),
            nn.Conv2d(c1, c1, kernel_size=3, stride=1, padding=autopad(3, None), dilation=dilation, groups=groups,
                      bias=False),
            nn.BatchNorm2d(c1),
        )
        self.k3 = nn.Sequential(
            nn.Conv2d(c1, c1, kernel_size=3, stride=1, padding=autopad(3, None), dilation=dilation, groups=groups,
                      bias=False),
            nn.BatchNorm2d(c1),
        )
        self.k4 = nn.Sequential(
            nn.Conv2d(c1, c2, kernel_size=3, stride=stride, padding=autopad(3, None), dilation=dilation, groups=groups,
                      bias=False),
            nn.BatchNorm2d(c2),
        )

    def forward(self, x):
        if kernel_path(self, x):
            return self.forward_b200(x)
        identity = x
        y_ = F.interpolate(self.k2(x), identity.size()[2:])
        y_ = torch.add(identity, y_)
        out = torch.sigmoid(y_)
        out = torch.mul(self.k3(x), out)
        return self.k4(out)

    def forward_b200(self, x, out=None):
        convs = (self.k2[1], self.k3[0], self.k4[0])
        if not all(_conv_supported(c) for c in convs):
            self._unsupported()
        x = ops.as_act(x)
        r = self.k2[0].kernel_size if isinstance(self.k2[0].kernel_size, int) else self.k2[0].kernel_size[0]
        pooled = getattr(x, '_dmay_pool4', None) if r == 4 else None     # by-product of the producing conv's epilogue
        if pooled is None or tuple(pooled.shape) != (x.shape[0], x.shape[1], x.shape[2] // 4, x.shape[3] // 4):
            pooled = ops.avgpool(x, r)
        k2o = ops.conv(pooled, get_conv_pack(self, 'k2', self.k2[1], self.k2[2], x.device), ACT_NONE)
        pk3 = get_conv_pack(self, 'k3', self.k3[0], self.k3[1], x.device)
        if _State.fuse_scconv_gate:
            g = ops.conv(x, pk3, ACT_NONE, gate=(x, k2o))       # k3(x) * sigmoid(x + up(k2)) in the epilogue
        else:
            g = ops.scconv_gate(x, ops.conv(x, pk3, ACT_NONE), k2o)
        return ops.conv(g, get_conv_pack(self, 'k4', self.k4[0], self.k4[1], x.device), ACT_NONE, out=out)

    def _unsupported(self):
        raise ops.DmayError("SCConv kernel path supports groups=1, dilation=1 only")


class space_to_depth(nn.Module):
    """SPD-Conv pixel-unshuffle — models/common.py:1451-1458."""

    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def forward(self, x):
        if kernel_path(self, x) and x.shape[1] % 8 == 0:
            x = _materialize(x)
            if ops.SPD_FOLD and x.dim() == 4 and x.shape[1] % 16 == 0 and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0:
                return ops.SPDView(ops.as_act(x))     # folded into a C3 consumer's cv1 | cv2, else materialised on demand
            return ops.spd(x)
        x = _materialize(x)
        return torch.cat([x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]], 1)


class SM(space_to_depth):
    """Same arithmetic under a second name — models/common.py:1460-1467."""
