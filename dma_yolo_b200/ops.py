"""Tensor-level wrappers over the C-ABI (include/dmayolo.h).

Activations are torch tensors with NCHW *shape* and channels_last *strides* (NHWC in memory),
bf16.  A channel slice `slab[:, a:b]` of such a tensor is also accepted everywhere ("ld" = the
slab's channel count); that is how concatenations are written in place by their producers.
PyTorch is used for device memory and streams only; every arithmetic op below is one of our kernels.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import torch

from . import _lib
from ._lib import CONSTS, DmayError, call

ACT_NONE, ACT_SILU, ACT_HSWISH, ACT_SIGMOID, ACT_GELU = 0, 1, 2, 3, 4
_DT = {torch.bfloat16: CONSTS["DMAY_DT_BF16"], torch.float32: CONSTS["DMAY_DT_F32"],
       torch.float16: CONSTS["DMAY_DT_F16"], torch.uint8: CONSTS["DMAY_DT_U8"]}


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


# --------------------------------------------------------------------------------------------
# layout helpers
# --------------------------------------------------------------------------------------------
def is_nhwc(t: torch.Tensor) -> bool:
    """NCHW-shaped tensor whose memory is [N,H,W,ld] (dense pixels, channel stride 1)."""
    if t.dim() != 4:
        return False
    n, c, h, w = t.shape
    sn, sc, sh, sw = t.stride()
    if c > 1 and sc != 1:
        return False
    ld = sw if w > 1 else (sh if h > 1 else (sn if n > 1 else c))
    if ld < c:
        return False
    if w > 1 and sw != ld:
        return False
    if h > 1 and sh != w * ld:
        return False
    if n > 1 and sn != h * w * ld:
        return False
    return True


def ld_of(t: torch.Tensor) -> int:
    n, c, h, w = t.shape
    sn, sc, sh, sw = t.stride()
    if w > 1:
        return sw
    if h > 1:
        return sh
    if n > 1:
        return sn
    return c


def empty_nhwc(n, c, h, w, device, dtype=torch.bfloat16, c_alloc=None) -> torch.Tensor:
    """NCHW-shaped, NHWC-in-memory tensor; c_alloc > c returns a channel-slice view of a padded slab."""
    ca = c_alloc or c
    buf = torch.empty((n, h, w, ca), device=device, dtype=dtype)
    t = buf.permute(0, 3, 1, 2)
    return t if ca == c else t[:, :c]


def as_act(x: torch.Tensor) -> torch.Tensor:
    """Any 4-D CUDA tensor -> bf16 NHWC activation (our layout kernel for contiguous NCHW input)."""
    if x.dtype == torch.bfloat16 and is_nhwc(x) and ld_of(x) % 8 == 0 and x.shape[1] % 8 == 0 and x.data_ptr() % 16 == 0:
        return x
    n, c, h, w = x.shape
    if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        x = x.float()
    if not x.is_contiguous():
        x = x.contiguous()
    cp = round_up(c, 8)
    if cp != c:
        y = torch.zeros((n, h, w, cp), device=x.device, dtype=torch.bfloat16).permute(0, 3, 1, 2)
    else:
        y = empty_nhwc(n, c, h, w, x.device)
    call("dmay_layout_convert", _stream(x), x=x.data_ptr(), y=y.data_ptr(), N=n, C=c, H=h, W=w, ld=cp,
         dtype=_DT[x.dtype], dir=1)
    return y if cp == c else y[:, :c]


def to_nchw(x: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """bf16 NHWC activation -> contiguous NCHW tensor of `dtype` (glue for torch-op modules)."""
    x = as_act(x)
    n, c, h, w = x.shape
    y = torch.empty((n, c, h, w), device=x.device, dtype=dtype)
    call("dmay_layout_convert", _stream(x), x=x.data_ptr(), y=y.data_ptr(), N=n, C=c, H=h, W=w, ld=ld_of(x),
         dtype=_DT[dtype], dir=0)
    return y


def input_prep(x: torch.Tensor, spd: bool, cpad: int = 16, mul: float = 1.0) -> torch.Tensor:
    """NCHW image {f32,f16,bf16,u8} -> NHWC bf16 [N,cpad,H(/2),W(/2)], optionally 2x2 pixel-unshuffled."""
    n, c, h, w = x.shape
    if not x.is_contiguous():
        x = x.contiguous()
    ho, wo = (h // 2, w // 2) if spd else (h, w)
    y = empty_nhwc(n, cpad, ho, wo, x.device)
    call("dmay_input_prep", _stream(x), x=x.data_ptr(), y=y.data_ptr(), N=n, C=c, H=h, W=w, Cpad=cpad,
         spd=int(spd), in_dtype=_DT[x.dtype], mul=float(mul))
    return y


# --------------------------------------------------------------------------------------------
# a1/a2 conv
# --------------------------------------------------------------------------------------------
@dataclass
class ConvPack:
    """Device-resident operands of one conv: bf16 weights [Cout_pad][kh][kw][Cin_pad] + fp32 scale/bias."""
    w: torch.Tensor
    scale: torch.Tensor
    bias: torch.Tensor
    cin: int          # logical input channels the caller passes
    cin_pad: int      # channels of the tensor the kernel reads
    cout: int
    cout_pad: int
    kh: int
    kw: int
    stride: int
    pad: int
    stem_spd: bool = False   # 2k x 2k stride-2 conv rewritten as k x k stride-1 over pixel-unshuffled input
    key: tuple = field(default_factory=tuple)


def pack_conv(weight: torch.Tensor, bn=None, conv_bias=None, stride=1, pad=0, device=None,
              allow_stem_spd=True) -> ConvPack:
    """Fold BN (utils/torch_utils.py:198-218 arithmetic, kept as fp32 per-channel scale/bias applied to the
    fp32 accumulator) and re-lay the weights K-major for the implicit GEMM."""
    device = device or weight.device
    w = weight.detach().float()
    cout, cin, kh, kw = w.shape
    if bn is not None:
        inv = torch.rsqrt(bn.running_var.detach().float() + bn.eps)
        g = bn.weight.detach().float() if bn.weight is not None else torch.ones_like(inv)
        b = bn.bias.detach().float() if bn.bias is not None else torch.zeros_like(inv)
        scale = g * inv
        bias = b - bn.running_mean.detach().float() * scale
        if conv_bias is not None:
            bias = bias + conv_bias.detach().float() * scale
    else:
        scale = torch.ones(cout, device=w.device)
        bias = conv_bias.detach().float() if conv_bias is not None else torch.zeros(cout, device=w.device)
    stem = (allow_stem_spd and cin <= 4 and kh == kw and kh % 2 == 0 and stride == 2 and pad % 2 == 0)
    wk = w.permute(0, 2, 3, 1)  # [Cout, kh, kw, Cin]
    if stem:
        k2 = kh // 2
        # W2[co, a, b, (dy + 2dx)*cin + c] = W[co, c, 2a+dy, 2b+dx]
        w6 = wk.reshape(cout, k2, 2, k2, 2, cin)          # co, a, dy, b, dx, c
        w2 = w6.permute(0, 1, 3, 4, 2, 5).reshape(cout, k2, k2, 4 * cin)  # (dx, dy, c) -> q = dy + 2dx
        wk, kh, kw, stride, pad = w2, k2, k2, 1, pad // 2
        cin_eff = 4 * cin
    else:
        cin_eff = cin
    cin_pad = round_up(cin_eff, 16)
    cout_pad = round_up(cout, 16)
    wp = torch.zeros((cout_pad, kh, kw, cin_pad), dtype=torch.float32, device=w.device)
    wp[:cout, :, :, :cin_eff] = wk
    sp = torch.zeros(cout_pad, dtype=torch.float32, device=w.device)
    bp = torch.zeros(cout_pad, dtype=torch.float32, device=w.device)
    sp[:cout] = scale
    bp[:cout] = bias
    return ConvPack(w=wp.to(device=device, dtype=torch.bfloat16).contiguous(), scale=sp.to(device).contiguous(),
                    bias=bp.to(device).contiguous(), cin=cin, cin_pad=cin_pad, cout=cout, cout_pad=cout_pad,
                    kh=kh, kw=kw, stride=stride, pad=pad, stem_spd=stem)


POOL4_FUSE = __import__('os').environ.get('DMAY_POOL4_FUSE', '1') != '0'   # A/B: 0 = SCConv always runs its AvgPool2d kernel
CONV_FLAGS = int(__import__('os').environ.get('DMAY_CONV_FLAGS', '0'))   # default tuning flags (see include/dmayolo.h)


def conv(x: torch.Tensor, pk: ConvPack, act: int = ACT_SILU, out: torch.Tensor | None = None,
         residual: torch.Tensor | None = None, gate: tuple | None = None, out_fp32: bool = False,
         block_n: int = 0, num_sms: int = 0, flags: int | None = None, res_mul: bool = False,
         pre: torch.Tensor | None = None, pool4: bool = False) -> torch.Tensor:
    """y = act(scale * conv(x, w) + bias) (+ residual) | (* sigmoid(gate_x + up(gate_k))).
    x may be a list of up to three tensors (virtual channel concat, 1x1 layers).  pre: fp32 NHWC [N, Cout, h, w] partial
    sums that are nearest-upsampled and added to the accumulator BEFORE scale / bias / SiLU (see VCat.split).
    pool4: also produce AvgPool2d(4)(y) in the epilogue where the kernel can (plain SiLU 3x3 layers with resident weights); the
    pooled tensor is attached to the result as `y._dmay_pool4` (absent when the layer does not qualify)."""
    # uint8 images are normalised on the fly (x/255, the `img.float()/255` of val.py:199-202 folded into
    # the layout kernel) — an extension: the reference only accepts float images.
    xs = None
    if isinstance(x, (list, tuple)):
        # virtual channel concat: the K loop of a 1x1 layer walks up to three tensors over the same pixels
        xs = [as_act(t) for t in x]
        x = xs[0]
        if not (pk.kh == 1 and pk.kw == 1 and pk.stride == 1 and pk.pad == 0) or pk.stem_spd or not 1 <= len(xs) <= 3:
            raise DmayError("conv: a list of inputs needs a 1x1 / stride-1 layer and at most three tensors")
        if any(t.shape[1] % 64 or t.shape[0] != x.shape[0] or t.shape[2:] != x.shape[2:] for t in xs):
            raise DmayError("conv: virtual concat parts need equal N, H, W and channel counts that are multiples of 64")
        if sum(t.shape[1] for t in xs) != pk.cin or pk.cin != pk.cin_pad:
            raise DmayError(f"conv: expected {pk.cin} input channels over all parts")
        if len(xs) == 1:
            xs = None
    mul = 1.0 / 255.0 if x.dtype == torch.uint8 else 1.0
    if xs is not None:
        pass
    elif pk.stem_spd:
        if x.shape[1] != pk.cin:
            raise DmayError(f"conv: expected {pk.cin} input channels, got {x.shape[1]}")
        x = input_prep(x, spd=True, cpad=pk.cin_pad, mul=mul)
    elif x.shape[1] != pk.cin_pad or not (x.dtype == torch.bfloat16 and is_nhwc(x)):
        if x.shape[1] != pk.cin:
            raise DmayError(f"conv: expected {pk.cin} input channels, got {x.shape[1]}")
        if pk.cin != pk.cin_pad:
            if x.dtype == torch.bfloat16 and is_nhwc(x):
                x = to_nchw(x, torch.bfloat16)
            x = input_prep(x, spd=False, cpad=pk.cin_pad, mul=mul)
        else:
            x = as_act(x if x.dtype != torch.uint8 else x.float() / 255)
    n, _, h, w = x.shape
    ho = (h + 2 * pk.pad - pk.kh) // pk.stride + 1
    wo = (w + 2 * pk.pad - pk.kw) // pk.stride + 1
    cstore = round_up(pk.cout, 8)
    odt = torch.float32 if out_fp32 else torch.bfloat16
    if out is None:
        out = empty_nhwc(n, pk.cout, ho, wo, x.device, odt, c_alloc=cstore)
    else:
        if tuple(out.shape) != (n, pk.cout, ho, wo) or out.dtype != odt or not is_nhwc(out):
            raise DmayError(f"conv: bad `out` {tuple(out.shape)} {out.dtype}, want {(n, pk.cout, ho, wo)} {odt} NHWC")
        if pk.cout % 8:
            raise DmayError("conv: writing into a slab slice needs Cout % 8 == 0")
    f = dict(x=x.data_ptr(), w=pk.w.data_ptr(), scale=pk.scale.data_ptr(), bias=pk.bias.data_ptr(), y=out.data_ptr(),
             N=n, H=h, W=w, Cin=pk.cin_pad, ldx=ld_of(x), Cout=cstore, Cout_pad=pk.cout_pad, kh=pk.kh, kw=pk.kw,
             stride=pk.stride, pad=pk.pad, Ho=ho, Wo=wo, ldy=ld_of(out), act=act,
             out_dtype=_DT[odt], block_n=block_n, num_sms=num_sms, flags=CONV_FLAGS if flags is None else flags)
    if xs is not None:
        f.update(x1=xs[1].data_ptr(), Cin1=xs[1].shape[1], ldx1=ld_of(xs[1]))
        if len(xs) == 3:
            f.update(x2=xs[2].data_ptr(), Cin2=xs[2].shape[1], ldx2=ld_of(xs[2]))
    if residual is not None:
        residual = as_act(residual)
        if tuple(residual.shape) != (n, pk.cout, ho, wo):
            raise DmayError("conv: residual shape mismatch")
        f.update(residual=residual.data_ptr(), ldr=ld_of(residual), res_op=int(res_mul))
    if pre is not None:
        if (pre.dtype != torch.float32 or not is_nhwc(pre) or pre.shape[0] != n or pre.shape[1] != pk.cout or act != ACT_SILU
                or out_fp32 or residual is not None or gate is not None):
            raise DmayError("conv: `pre` needs fp32 NHWC partial sums of the output's channels and a plain SiLU bf16 layer")
        f.update(pre=pre.data_ptr(), ldpre=ld_of(pre), preH=pre.shape[2], preW=pre.shape[3])
    if gate is not None:
        gx, gk = gate
        gx, gk = as_act(gx), as_act(gk)
        if tuple(gx.shape) != (n, pk.cout, ho, wo) or gk.shape[1] != pk.cout or ld_of(gk) != pk.cout:
            raise DmayError("conv: gate shape mismatch")
        f.update(gate_x=gx.data_ptr(), gate_k=gk.data_ptr(), ldgx=ld_of(gx), gHk=gk.shape[2], gWk=gk.shape[3])
    if pool4 and POOL4_FUSE and act == ACT_SILU and not out_fp32 and ho % 4 == 0 and wo % 4 == 0 and pk.kh == 3 and pk.stride == 1:
        pooled = empty_nhwc(n, pk.cout, ho // 4, wo // 4, x.device)
        try:
            call("dmay_conv_bn_act", _stream(x), pool4_out=pooled.data_ptr(), ldpool4=ld_of(pooled), **f)
            out._dmay_pool4 = pooled
            return out
        except DmayError:       # not a layer whose kernel image carries the pooling epilogue: plain launch below
            pass
    call("dmay_conv_bn_act", _stream(x), **f)
    return out


# --------------------------------------------------------------------------------------------
# memory-bound blocks
# --------------------------------------------------------------------------------------------
def spd(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    x = as_act(x)
    n, c, h, w = x.shape
    if out is None:
        out = empty_nhwc(n, 4 * c, h // 2, w // 2, x.device)
    call("dmay_spd", _stream(x), x=x.data_ptr(), y=out.data_ptr(), N=n, H=h, W=w, C=c, ldx=ld_of(x), ldy=ld_of(out))
    return out


class Up:
    """A nearest-neighbour 2^k upsample that has not been materialised (fused into its consumer)."""

    def __init__(self, src: torch.Tensor, log2f: int):
        self.src, self.log2f = src, log2f

    @property
    def shape(self):
        n, c, h, w = self.src.shape
        return torch.Size((n, c, h << self.log2f, w << self.log2f))

    def materialize(self) -> torch.Tensor:
        return upsample(self.src, 1 << self.log2f)


class VCat(Up):
    """cat([w_i * x_i], 1) that has not been materialised (AdConcat2 / AdConcat3 / Concat outputs).  A 1x1 consumer (C3's
    cv1 | cv2, Conv) reads the parts in place -- `sources()`, the K loop of its GEMM walks them -- with w_i folded into the
    matching columns of its weights (`colscale()`); every other consumer gets `materialize()` (the adconcat kernel, once)."""

    def __init__(self, parts, weights):
        self.parts = list(parts)
        self.weights = tuple(float(w) for w in weights)
        first = self.parts[0]
        self.src, self.log2f = (first.src if isinstance(first, Up) else first), 0
        self._mat = self._srcs = None

    @property
    def shape(self):
        n, _, h, w = self.parts[0].shape
        return torch.Size((n, sum(p.shape[1] for p in self.parts), h, w))

    def materialize(self) -> torch.Tensor:
        if self._mat is None:
            self._mat = adconcat(self.parts, self.weights)
        return self._mat

    def sources(self):
        """Same-resolution bf16 NHWC tensors for a virtual-concat GEMM (an `Up` part is replicated once), or None."""
        if self._srcs is None:
            if any(p.shape[1] % 64 for p in self.parts):
                return None
            self._srcs = [as_act(p.materialize() if isinstance(p, Up) else p) for p in self.parts]
        return self._srcs

    def colscale(self):
        """((channels, weight), ...) per part; None when every weight is exactly 1 (plain Concat: nothing to fold)."""
        if all(w == 1.0 for w in self.weights):
            return None
        return tuple((int(p.shape[1]), w) for p, w in zip(self.parts, self.weights))

    def split(self):
        """A 1x1 layer commutes with a nearest upsample: W.[up(x0) | x1] = up(W0.x0) + W1.x1.  -> (low, high, cols) with
        low = the sources of the `Up` parts (all at ONE reduced resolution), high = the same-resolution parts, cols =
        ((start, stop) column ranges of the low parts, ... of the high parts); None when the concat has no such split
        (no `Up` part, nothing but `Up` parts, mixed factors, or channel counts that are not multiples of 64)."""
        if not SPLIT_UP or any(p.shape[1] % 64 for p in self.parts):
            return None
        ups = [p for p in self.parts if isinstance(p, Up)]
        if not ups or len(ups) == len(self.parts) or len({p.log2f for p in ups}) != 1 or any(isinstance(p, VCat) for p in ups):
            return None
        lo, hi, lo_cols, hi_cols, c0 = [], [], [], [], 0
        for p in self.parts:
            c = int(p.shape[1])
            if isinstance(p, Up):
                lo.append(as_act(p.src))
                lo_cols.append((c0, c0 + c))
            else:
                hi.append(as_act(p))
                hi_cols.append((c0, c0 + c))
            c0 += c
        return lo, hi, (tuple(lo_cols), tuple(hi_cols))


SPD_FOLD = __import__('os').environ.get('DMAY_SPD_FOLD', '1') != '0'   # A/B switch: 0 = space_to_depth always runs its kernel


class SPDView(Up):
    """A space_to_depth (SPD-Conv pixel unshuffle, models/common.py:1451-1458) that has not been materialised.  A 1x1 layer over
    its output IS a 2x2 / stride-2 convolution over its input -- y[co] = sum_{dy,dx,c} W[co, (dy + 2 dx) C + c] x[2h+dy, 2w+dx, c]
    -- so C3's cv1 | cv2 read `src` through the im2col maps with re-laid weights and the 4C-channel tensor is never written;
    every other consumer gets `materialize()` (dmay_spd)."""

    def __init__(self, src: torch.Tensor):
        self.src, self.log2f = src, 0
        self._mat = None

    @property
    def shape(self):
        n, c, h, w = self.src.shape
        return torch.Size((n, 4 * c, h // 2, w // 2))

    def materialize(self) -> torch.Tensor:
        if self._mat is None:
            self._mat = spd(self.src)
        return self._mat


def adconcat(xs, weights, out: torch.Tensor | None = None) -> torch.Tensor:
    """cat([w_i * x_i], 1); x_i may be an `Up` (read at reduced resolution)."""
    if not 2 <= len(xs) <= 3:
        raise DmayError("adconcat takes 2 or 3 inputs")
    srcs, ups = [], []
    for t in xs:
        if isinstance(t, (VCat, SPDView)):
            t = t.materialize()
        if isinstance(t, Up):
            srcs.append(as_act(t.src))
            ups.append(t.log2f)
        else:
            srcs.append(as_act(t))
            ups.append(0)
    n, _, h0, w0 = srcs[0].shape
    h, w = h0 << ups[0], w0 << ups[0]
    for s, u in zip(srcs, ups):
        if (s.shape[2] << u, s.shape[3] << u) != (h, w) or s.shape[0] != n:
            raise DmayError("adconcat: spatial/batch mismatch")
    ctot = sum(s.shape[1] for s in srcs)
    if out is None:
        out = empty_nhwc(n, ctot, h, w, srcs[0].device)
    f = dict(y=out.data_ptr(), n_in=len(srcs), N=n, H=h, W=w, ldy=ld_of(out))
    for i, (s, u, wt) in enumerate(zip(srcs, ups, weights)):
        f[f"x{i}"] = s.data_ptr()
        f[f"C{i}"] = s.shape[1]
        f[f"ld{i}"] = ld_of(s)
        f[f"up{i}"] = u
        f[f"w{i}"] = float(wt)
    call("dmay_adconcat", _stream(srcs[0]), **f)
    return out


def concat(xs, out=None):
    """torch.cat(xs, 1) for 2..3 inputs (weights 1.0: an exact copy); longer lists are folded pairwise."""
    xs = [t.materialize() if isinstance(t, (VCat, SPDView)) else t for t in xs]
    while len(xs) > 3:
        xs = [adconcat(xs[:3], (1.0, 1.0, 1.0))] + xs[3:]
    if len(xs) == 1:
        return xs[0].materialize() if isinstance(xs[0], Up) else xs[0]
    return adconcat(xs, (1.0,) * len(xs), out=out)


VCAT = __import__('os').environ.get('DMAY_VCAT', '1') != '0'   # A/B switch: 0 = every concat is materialised (round-1 form)
# 1: the `Up` parts of a lazy concat are split off into a low-resolution partial GEMM whose fp32 sums the main GEMM's epilogue adds
# (VCat.split, EPI_SILU_PRE).  Measured equal to replicating the part (same box, cfg-2: 12.87 vs 12.84 ms per step -- the two
# consumers are L2 / HBM-bound 1x1 layers, the saved flops do not show and the fp32 partials cost what the replicate did): opt-in.
SPLIT_UP = __import__('os').environ.get('DMAY_SPLIT_UP', '0') == '1'


def vcat(xs, weights):
    """Lazy cat([w_i * x_i], 1) when a 1x1 consumer could read the parts in place, else the materialised tensor."""
    xs = [t.materialize() if isinstance(t, (VCat, SPDView)) else t for t in xs]
    if VCAT and 2 <= len(xs) <= 3 and all(t.shape[1] % 64 == 0 for t in xs):
        return VCat(xs, weights)
    return adconcat(xs, weights) if len(xs) <= 3 else concat(xs)


def adapt_add(xs, weights, out=None) -> torch.Tensor:
    xs = [as_act(t) for t in xs]
    n, c, h, w = xs[0].shape
    if out is None:
        out = empty_nhwc(n, c, h, w, xs[0].device)
    f = dict(y=out.data_ptr(), n_in=len(xs), npix=n * h * w, C=c, ldy=ld_of(out))
    for i, (s, wt) in enumerate(zip(xs, weights)):
        f[f"x{i}"] = s.data_ptr()
        f[f"ld{i}"] = ld_of(s)
        f[f"w{i}"] = float(wt)
    call("dmay_adaptadd", _stream(xs[0]), **f)
    return out


def upsample(x: torch.Tensor, factor: int = 2, out=None) -> torch.Tensor:
    x = as_act(x)
    n, c, h, w = x.shape
    if out is None:
        out = empty_nhwc(n, c, h * factor, w * factor, x.device)
    call("dmay_upsample_nearest", _stream(x), x=x.data_ptr(), y=out.data_ptr(), N=n, H=h, W=w, C=c, ldx=ld_of(x),
         ldy=ld_of(out), factor=factor)
    return out


def sppf_pool3(x: torch.Tensor, y1, y2, y3, k: int = 5):
    n, c, h, w = x.shape
    call("dmay_sppf_pool3", _stream(x), x=x.data_ptr(), y1=y1.data_ptr(), y2=y2.data_ptr(), y3=y3.data_ptr(), N=n, H=h,
         W=w, C=c, ldx=ld_of(x), ldy=ld_of(y1), k=k)


def maxpool_s1(x: torch.Tensor, k: int, out=None) -> torch.Tensor:
    x = as_act(x)
    n, c, h, w = x.shape
    if out is None:
        out = empty_nhwc(n, c, h, w, x.device)
    call("dmay_maxpool_s1", _stream(x), x=x.data_ptr(), y=out.data_ptr(), N=n, H=h, W=w, C=c, ldx=ld_of(x),
         ldy=ld_of(out), k=k)
    return out


def avgpool(x: torch.Tensor, r: int, out=None) -> torch.Tensor:
    x = as_act(x)
    n, c, h, w = x.shape
    if out is None:
        out = empty_nhwc(n, c, h // r, w // r, x.device)
    call("dmay_avgpool", _stream(x), x=x.data_ptr(), y=out.data_ptr(), N=n, H=h, W=w, C=c, ldx=ld_of(x), ldy=ld_of(out),
         r=r)
    return out


def scconv_gate(x, k3, k2, out=None) -> torch.Tensor:
    x, k3, k2 = as_act(x), as_act(k3), as_act(k2)
    n, c, h, w = x.shape
    if out is None:
        out = empty_nhwc(n, c, h, w, x.device)
    call("dmay_scconv_gate", _stream(x), x=x.data_ptr(), k3=k3.data_ptr(), k2=k2.data_ptr(), y=out.data_ptr(), N=n, H=h,
         W=w, C=c, Hk=k2.shape[2], Wk=k2.shape[3], ldx=ld_of(x), ld3=ld_of(k3), ld2=ld_of(k2), ldy=ld_of(out))
    return out


@dataclass
class CoordAttPack:
    w1: torch.Tensor
    b1: torch.Tensor
    s1: torch.Tensor
    t1: torch.Tensor
    whT: torch.Tensor
    bh: torch.Tensor
    wwT: torch.Tensor
    bw: torch.Tensor
    c: int
    cm: int
    key: tuple = field(default_factory=tuple)


def pack_coordatt(conv1, bn1, conv_h, conv_w, device) -> CoordAttPack:
    """models/common.py:1168-1181 parameters -> fp32 device operands (BN folded to s1/t1; all three weight
    matrices transposed: w1 [C][Cm], whT/wwT [Cm][Cout])."""
    cm, c = conv1.weight.shape[:2]
    f = lambda t: t.detach().float().to(device).contiguous()
    inv = torch.rsqrt(bn1.running_var.detach().float() + bn1.eps)
    s1 = bn1.weight.detach().float() * inv
    t1 = bn1.bias.detach().float() - bn1.running_mean.detach().float() * s1
    b1 = conv1.bias if conv1.bias is not None else torch.zeros(cm)
    zeros = lambda m: m.bias if m.bias is not None else torch.zeros(m.weight.shape[0])
    return CoordAttPack(w1=f(conv1.weight.reshape(cm, c).t()), b1=f(b1), s1=f(s1), t1=f(t1),
                        whT=f(conv_h.weight.reshape(-1, cm).t()), bh=f(zeros(conv_h)),
                        wwT=f(conv_w.weight.reshape(-1, cm).t()), bw=f(zeros(conv_w)), c=c, cm=cm)


_CA_WS: dict = {}   # (device, N, H, W, C, Cm) -> zero-initialised workspace (arrival tickets reset themselves)


def coordatt(x: torch.Tensor, pk: CoordAttPack, out=None, return_gates=False):
    x = as_act(x)
    n, c, h, w = x.shape
    if c != pk.c or pk.whT.shape[1] != c:
        raise DmayError("coordatt: channel mismatch (needs c2 == c1)")
    if out is None:
        out = empty_nhwc(n, c, h, w, x.device)
    key = (x.device.index, n, h, w, c, pk.cm, torch.cuda.current_stream(x.device).cuda_stream)
    ws = _CA_WS.get(key)
    if ws is None:
        nbytes = int(_lib.lib().dmay_coordatt_ws(n, h, w, c, pk.cm))
        ws = _CA_WS[key] = torch.zeros((nbytes + 15) // 16 * 4, device=x.device, dtype=torch.int32)
    f = dict(x=x.data_ptr(), y=out.data_ptr(), w1=pk.w1.data_ptr(), b1=pk.b1.data_ptr(), s1=pk.s1.data_ptr(),
             t1=pk.t1.data_ptr(), wh=pk.whT.data_ptr(), bh=pk.bh.data_ptr(), ww=pk.wwT.data_ptr(), bw=pk.bw.data_ptr(),
             N=n, H=h, W=w, C=c, Cm=pk.cm, ldx=ld_of(x), ldy=ld_of(out), num_sms=0, ws=ws.data_ptr(), ws_bytes=ws.numel() * 4)
    pooled = gates = None
    plane_fits = (h * w * 128 + (h + w) * 256 + 256 * pk.cm <= 100 * 1024
                  and ((h + w) * (pk.cm + 64) + 128 * pk.cm) * 4 <= 100 * 1024 and pk.cm <= 256)   # mirrors coordatt.cu
    if return_gates or not plane_fits:   # the three-launch path (large planes) needs both; tests ask for them
        pooled = torch.empty((n, h + w, c), device=x.device, dtype=torch.float32)
        gates = torch.empty((n, h + w, c), device=x.device, dtype=torch.float32)
        f.update(pooled=pooled.data_ptr(), gates=gates.data_ptr())
    call("dmay_coordatt", _stream(x), **f)
    return (out, pooled, gates) if return_gates else out


# --------------------------------------------------------------------------------------------
# 8f-1 Swin pieces (C3STR): LayerNorm over channels, window attention
# --------------------------------------------------------------------------------------------
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, out=None) -> torch.Tensor:
    """nn.LayerNorm(C) applied to the channel vector of every pixel of an NHWC activation."""
    x = as_act(x)
    n, c, h, w = x.shape
    if out is None:
        out = empty_nhwc(n, c, h, w, x.device)
    call("dmay_layernorm", _stream(x), x=x.data_ptr(), y=out.data_ptr(), gamma=gamma.data_ptr(), beta=beta.data_ptr(),
         npix=n * h * w, C=c, ldx=ld_of(x), ldy=ld_of(out), eps=float(eps))
    return out


def window_attention(qkv: torch.Tensor, rel_bias: torch.Tensor, mask, heads: int, shift: int, scale: float, out=None,
                     variant: int = 0):
    """W-MSA / SW-MSA on the qkv tensor [N, 3C, H, W] (NHWC) -> [N, C, H, W]; padding, shift and window (un)partition are
    index arithmetic inside the kernel (models/common.py:483-515, 603-627)."""
    qkv = as_act(qkv)
    n, c3, h, w = qkv.shape
    c = c3 // 3
    if out is None:
        out = empty_nhwc(n, c, h, w, qkv.device)
    call("dmay_window_attention", _stream(qkv), qkv=qkv.data_ptr(), out=out.data_ptr(), rel_bias=rel_bias.data_ptr(),
         mask=mask.data_ptr() if mask is not None else 0, N=n, H=h, W=w, C=c, heads=heads, window=8, shift=int(shift),
         ldq=ld_of(qkv), ldo=ld_of(out), scale=float(scale), variant=int(variant))
    return out


# --------------------------------------------------------------------------------------------
# a8 Detect decode + a9 NMS
# --------------------------------------------------------------------------------------------
@dataclass
class DetectLevel:
    logits: torch.Tensor      # fp32 [N, ny, nx, ld] contiguous (NHWC), channel = a*no + o
    stride: float
    anchors_px: list          # [(w, h)] * na, pixels (= anchors * stride)
    ny: int
    nx: int
    ld: int
    pitch: int = 0            # floats per anchor row inside a pixel's channel vector (0 = no; 4n when the head pads the rows)


def _level_meta_host(levels, na):
    words, row0 = [], 0
    for lv in levels:
        flat = [float(v) for wh in lv.anchors_px for v in wh] + [0.0] * (10 - 2 * na)
        words.append(struct.pack("<5if10f", row0, lv.ny, lv.nx, lv.ld, na, float(lv.stride), *flat))
        row0 += na * lv.ny * lv.nx
    return b"".join(words), row0


def _level_meta(levels, na, device):
    raw, row0 = _level_meta_host(levels, na)
    buf = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
    return buf, row0


# candidate-buffer capacity of the fused filter, remembered per (device, batch, rows, nc, multi_label): the
# kernel counts every candidate but only writes those below `capacity`, so an undersized guess costs one re-run.
_FUSED_CAP: dict = {}
FUSED_FILTER = __import__('os').environ.get('DMAY_FUSED_FILTER', '1') != '0'


def detect_decode(levels, na: int, no: int) -> torch.Tensor:
    """Dense `pred` [N, sum(na*ny*nx), no] fp32 exactly as Detect.forward returns it (models/yolo.py:81-101)."""
    n = levels[0].logits.shape[0]
    rows = sum(na * lv.ny * lv.nx for lv in levels)
    pred = torch.empty((n, rows, no), device=levels[0].logits.device, dtype=torch.float32)
    row0 = 0
    for lv in levels:
        a = [v for wh in lv.anchors_px for v in wh] + [0.0] * (10 - 2 * na)
        call("dmay_detect_decode", _stream(pred), logits=lv.logits.data_ptr(), pred=pred.data_ptr(), N=n, ny=lv.ny,
             nx=lv.nx, na=na, no=no, ld=lv.ld, row_pitch=lv.pitch, row0=row0, rows_total=rows, stride=float(lv.stride),
             aw0=a[0], ah0=a[1], aw1=a[2], ah1=a[3], aw2=a[4], ah2=a[5], aw3=a[6], ah3=a[7], aw4=a[8], ah4=a[9])
        row0 += na * lv.ny * lv.nx
    return pred


def _det_buffers(n, max_det, dev):
    """One zeroed fp32 buffer [N*max_det*6 | N counts (int32 bits)] and its two views: the layout the one-collective
    all-gather of dist.py moves as is."""
    buf = torch.zeros(n * max_det * 6 + n, device=dev, dtype=torch.float32)
    return buf, buf[:n * max_det * 6].view(n, max_det, 6), buf[n * max_det * 6:].view(torch.int32)


def nms_batched(pred: torch.Tensor | None, conf_thres: float, iou_thres: float, *, levels=None, na=0, nc=None,
                classes=None, agnostic=False, multi_label=False, max_det=300, max_nms=30000, max_wh=4096.0,
                return_packed=False):
    """Whole-batch NMS.  Source is a dense fp32 `pred` [N,R,5+nc] or Detect `levels` (fused decode).
    Returns (out [N,max_det,6] fp32, counts [N] int32) on the device; out[i,:counts[i]] are image i's
    detections (xyxy, conf, cls) in descending score order — utils/general.py:633-725.  `return_packed` appends the
    flat buffer both are views of."""
    r = _nms_batched(pred, conf_thres, iou_thres, levels, na, nc, classes, agnostic, multi_label, max_det, max_nms, max_wh)
    return r if return_packed else r[:2]


def _nms_batched(pred, conf_thres, iou_thres, levels, na, nc, classes, agnostic, multi_label, max_det, max_nms, max_wh):
    if levels is not None:
        dev = levels[0].logits.device
        n = levels[0].logits.shape[0]
        if FUSED_FILTER:
            return _nms_fused(levels, na, nc, conf_thres, iou_thres, classes, agnostic, multi_label, max_det, max_nms, max_wh)
        meta, rows = _level_meta(levels, na, dev)
        src = dict(levels=len(levels), lv_meta=meta.data_ptr(), row_pitch=levels[0].pitch)
        for i, lv in enumerate(levels):
            src[f"lv_logits{i}"] = lv.logits.data_ptr()
        keep_alive = (meta,)
    else:
        if pred.dtype != torch.float32 or not pred.is_contiguous():
            pred = pred.float().contiguous()
        dev = pred.device
        n, rows, no = pred.shape
        nc = no - 5
        src = dict(levels=0, pred=pred.data_ptr())
        keep_alive = (pred,)
    s = torch.cuda.current_stream(dev).cuda_stream
    multi_label = bool(multi_label) and nc > 1
    buf, out, out_counts = _det_buffers(n, max_det, dev)
    if levels is None and DENSE_ROWS_FILTER and nc <= 96:
        # dense prediction through the thread-per-row single-pass filter (one read of the tensor instead of the
        # count + write passes of the warp-per-row kernels)
        keys, cand, img_counts, img_offsets, offs_host, total = _fused_candidates(_dense_level(pred), 1, nc, conf_thres,
                                                                                  multi_label, classes, dense=True, max_nms=max_nms)
    else:
        keys, cand, img_counts, img_offsets, offs_host, total = _ordered_candidates(src, n, rows, nc, conf_thres, multi_label,
                                                                                    classes, dev, s)
    if total == 0:
        return out, out_counts, buf
    idx, cnts, offs = _order_candidates(keys, offs_host, n, img_counts, img_offsets, max_nms, dev, s)
    call("dmay_nms_greedy", s, cand=cand.data_ptr(), sorted_idx=idx.data_ptr(), img_counts=cnts.data_ptr(),
         img_offsets=offs.data_ptr(), out=out.data_ptr(), out_counts=out_counts.data_ptr(), N=n, max_det=max_det,
         max_nms=max_nms, agnostic=int(bool(agnostic)), max_wh=float(max_wh), iou_thres=float(iou_thres))
    del keep_alive
    return out, out_counts, buf


DENSE_ROWS_FILTER = __import__('os').environ.get('DMAY_DENSE_ROWS', '1') != '0'
_DENSE_RESERVE = __import__('os').environ.get('DMAY_DENSE_RESERVE')


def _dense_level(pred: torch.Tensor):
    """A dense prediction [N, R, no] described as ONE pseudo level (na = 1, ny = 1, nx = R, ld = no) for the fused filter."""
    n, rows, no = pred.shape
    return [DetectLevel(logits=pred, stride=0.0, anchors_px=[(0.0, 0.0)], ny=1, nx=rows, ld=no, pitch=0)]


def _ordered_candidates(src, n, rows, nc, conf_thres, multi_label, classes, dev, s):
    """Three-launch order-preserving filter (count / scan / write) of a dense prediction or of Detect levels."""
    rpb = 128
    nblk = (rows + rpb - 1) // rpb
    blk_counts = torch.empty(n * nblk, device=dev, dtype=torch.int32)
    blk_offsets = torch.empty(n * nblk, device=dev, dtype=torch.int32)
    img_counts = torch.empty(n, device=dev, dtype=torch.int32)
    img_offsets = torch.empty(n + 1, device=dev, dtype=torch.int64)
    common = dict(blk_counts=blk_counts.data_ptr(), blk_offsets=blk_offsets.data_ptr(), img_counts=img_counts.data_ptr(),
                  img_offsets=img_offsets.data_ptr(), N=n, R=rows, nc=nc, multi_label=int(multi_label),
                  rows_per_block=rpb, conf_thres=float(conf_thres), **src)
    cm = _class_mask(classes, nc, dev)
    if cm is not None:
        common["class_mask"] = cm.data_ptr()
    call("dmay_nms_filter", s, phase=0, **common)
    offs_host = img_offsets.tolist()    # the one sizing sync of the batch (N+1 values)
    total = offs_host[-1]
    if total == 0:
        return None, None, img_counts, img_offsets, offs_host, 0
    keys = torch.empty(total, device=dev, dtype=torch.int64)
    cand = torch.empty((total, 6), device=dev, dtype=torch.float32)
    call("dmay_nms_filter", s, phase=1, keys=keys.data_ptr(), cand=cand.data_ptr(), capacity=total, **common)
    del cm
    return keys, cand, img_counts, img_offsets, offs_host, total


def _class_mask(classes, nc, dev):
    if classes is None:
        return None
    cm = torch.zeros(nc, dtype=torch.uint8)
    for c in classes:
        if 0 <= int(c) < nc:
            cm[int(c)] = 1
    return cm.to(dev)


def _sort_candidates(keys, total, n, dev, s, vals=None):
    """Stable radix sort of the (image, ~score) keys; payload = candidate index (`vals`, or the position in `keys`)."""
    img_bits = max(1, (n - 1).bit_length())
    ws_bytes = int(_lib.lib().dmay_nms_sort_ws(total, img_bits))
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    keys_out = torch.empty(total, device=dev, dtype=torch.int64)
    idx = torch.empty(total, device=dev, dtype=torch.int32)
    f = dict(keys_in=keys.data_ptr(), keys_out=keys_out.data_ptr(), idx_out=idx.data_ptr(), ws=ws.data_ptr(),
             ws_bytes=ws_bytes, n=total, img_bits=img_bits)
    if vals is not None:
        f["vals_in"] = vals.data_ptr()
    call("dmay_nms_sort", s, **f)
    return keys_out, idx


TOPK_SELECT = __import__('os').environ.get('DMAY_TOPK_SELECT', '1') != '0'


def _order_candidates(keys, offs_host, n, img_counts, img_offsets, max_nms, dev, s):
    """-> (sorted candidate indices, per-image counts, per-image offsets into them) for the greedy kernel.
    When an image has more than max_nms candidates, the exact top-max_nms pre-selection runs first and only the
    survivors are sorted (utils/general.py:702-703 keeps just those anyway)."""
    total = offs_host[-1]
    counts = [offs_host[i + 1] - offs_host[i] for i in range(n)]
    if not TOPK_SELECT or max(counts) <= max_nms:
        _ko, idx = _sort_candidates(keys, total, n, dev, s)
        return idx, img_counts, img_offsets
    n_c = sum(min(c, max_nms) for c in counts)
    keys_c = torch.empty(n_c, device=dev, dtype=torch.int64)
    idx_c = torch.empty(n_c, device=dev, dtype=torch.int32)
    counts_c = torch.empty(n, device=dev, dtype=torch.int32)
    offs_c = torch.empty(n + 1, device=dev, dtype=torch.int64)
    call("dmay_nms_topk_select", s, keys=keys.data_ptr(), img_counts=img_counts.data_ptr(), img_offsets=img_offsets.data_ptr(),
         keys_out=keys_c.data_ptr(), idx_out=idx_c.data_ptr(), counts_out=counts_c.data_ptr(), offsets_out=offs_c.data_ptr(),
         N=n, K=max_nms)
    _ko, idx = _sort_candidates(keys_c, n_c, n, dev, s, vals=idx_c)
    return idx, counts_c, offs_c


def _nms_fused(levels, na, nc, conf_thres, iou_thres, classes, agnostic, multi_label, max_det, max_nms, max_wh):
    """Detect-logits source: single-pass fused decode + filter + order-preserving compaction -> sort -> greedy."""
    import ctypes
    dev = levels[0].logits.device
    n = levels[0].logits.shape[0]
    s = torch.cuda.current_stream(dev).cuda_stream
    multi_label = bool(multi_label) and nc > 1
    buf, out, out_counts = _det_buffers(n, max_det, dev)
    keys, cand, img_counts, img_offsets, offs_host, total = _fused_candidates(levels, na, nc, conf_thres, multi_label, classes,
                                                                             max_nms=max_nms)
    if total == 0:
        return out, out_counts, buf
    idx, cnts, offs = _order_candidates(keys, offs_host, n, img_counts, img_offsets, max_nms, dev, s)
    call("dmay_nms_greedy", s, cand=cand.data_ptr(), sorted_idx=idx.data_ptr(), img_counts=cnts.data_ptr(),
         img_offsets=offs.data_ptr(), out=out.data_ptr(), out_counts=out_counts.data_ptr(), N=n, max_det=max_det,
         max_nms=max_nms, agnostic=int(bool(agnostic)), max_wh=float(max_wh), iou_thres=float(iou_thres))
    return out, out_counts, buf


PRETHRESHOLD = __import__('os').environ.get('DMAY_PRETHRESHOLD', '1') != '0'
# Detect-logits source: the histogram pre-selection costs one more pass over the logits, so it is switched on per (shape, threshold)
# once a call has seen more than this many times max_nms candidates per image ON AVERAGE (cfg-4a / cfg-4b: 1.7 M per image; cfg-2
# has single images above 120 k but pays 0.11 ms per step for the extra pass: same box 12.50 vs 12.62 ms), 0 = never
FUSED_PRETHR_RATIO = float(__import__('os').environ.get('DMAY_FUSED_PRETHR_RATIO', '4'))
_FUSED_PRETHR = {}
PER_IMAGE_REGIONS = __import__('os').environ.get('DMAY_FILTER_REGIONS', '1') != '0'   # one reservation counter per image


def _fused_candidates(levels, na, nc, conf_thres, multi_label, classes, dense=False, reserve=None, max_nms=0):
    """Single-pass fused decode + filter + order-preserving compaction of the Detect logits (multi_label already
    reduced by `nc > 1`).  `dense`: `levels` is one pseudo level wrapping a dense prediction [N, R, 5 + nc] (values used
    as they are).  -> keys, cand, img_counts, img_offsets, img_offsets as a host list, total."""
    import ctypes
    dev = levels[0].logits.device
    n = levels[0].logits.shape[0]
    s = torch.cuda.current_stream(dev).cuda_stream
    raw, rows = _level_meta_host(levels, na)
    meta_host = ctypes.create_string_buffer(raw, len(raw))
    ws_bytes = int(_lib.lib().dmay_nms_filter_fused_ws(ctypes.addressof(meta_host), len(levels), n))
    if ws_bytes < 0:
        raise DmayError(f"dmay_nms_filter_fused_ws failed: {ws_bytes}")
    cm = _class_mask(classes, nc, dev)
    key = (dev.index, n, rows, nc, multi_label, float(conf_thres), bool(dense))
    capacity = _FUSED_CAP.get(key, n * rows * (2 if multi_label else 1) // 2 + 4096)
    bin_thr = None
    if dense and multi_label and max_nms > 0 and PRETHRESHOLD and rows * nc > max_nms:
        # an image may hold more than max_nms candidates: only those that the top-max_nms selection could keep are written
        hist = torch.zeros((n, 2048), device=dev, dtype=torch.int32)
        bin_thr = torch.empty(n, device=dev, dtype=torch.int32)
        pf = dict(pred=levels[0].logits.data_ptr(), hist=hist.data_ptr(), bin_thr=bin_thr.data_ptr(), N=n, R=rows, nc=nc,
                  K=int(max_nms), conf_thres=float(conf_thres))
        if cm is not None:
            pf["class_mask"] = cm.data_ptr()
        call("dmay_nms_dense_prethreshold", s, **pf)
    img_counts = torch.empty(n, device=dev, dtype=torch.int32)
    img_offsets = torch.empty(n + 1, device=dev, dtype=torch.int64)
    fused_pre = (not dense and multi_label and max_nms > 0 and PRETHRESHOLD and _FUSED_PRETHR.get(key, False)
                 and levels[0].pitch > 5 + nc and nc <= 96)
    if fused_pre:
        hist = torch.zeros((n, 2048), device=dev, dtype=torch.int32)
        bin_thr = torch.empty(n, device=dev, dtype=torch.int32)
        pf = dict(lv_meta_host=ctypes.addressof(meta_host), N=n, nc=nc, levels=len(levels), multi_label=1,
                  conf_thres=float(conf_thres), row_pitch=levels[0].pitch, hist=hist.data_ptr(), bin_thr=bin_thr.data_ptr(),
                  prethr_k=int(max_nms))
        for i, lv in enumerate(levels):
            pf[f"lv_logits{i}"] = lv.logits.data_ptr()
        if cm is not None:
            pf["class_mask"] = cm.data_ptr()
        call("dmay_nms_fused_prethreshold", s, **pf)
    while True:
        ws = torch.zeros(ws_bytes // 8 + 1, device=dev, dtype=torch.int64)   # ticket + tile status words (zeroed)
        keys = torch.empty(capacity, device=dev, dtype=torch.int64)
        cand = torch.empty((capacity, 6), device=dev, dtype=torch.float32)
        f = dict(lv_meta_host=ctypes.addressof(meta_host), ws=ws.data_ptr(), ws_bytes=ws.numel() * 8,
                 img_counts=img_counts.data_ptr(), img_offsets=img_offsets.data_ptr(), keys=keys.data_ptr(),
                 cand=cand.data_ptr(), N=n, nc=nc, levels=len(levels), multi_label=int(multi_label), capacity=capacity,
                 conf_thres=float(conf_thres), row_pitch=levels[0].pitch, dense=int(dense))
        # temporary buffers of the reserve + scan + gather mode (tiles place their runs without waiting for each other).
        # It moves every candidate twice more (32 B each way), so it is for sparse outputs: where most (row, class) pairs
        # become candidates (dense multi-label predictions: cfg-5 writes 2 GB of candidates for 0.39 GB of input) the
        # single-launch look-back form writes them once, in place.
        if bin_thr is not None:
            f["bin_thr"] = bin_thr.data_ptr()
        if dense and _DENSE_RESERVE is not None and reserve is None:
            reserve = _DENSE_RESERVE == '1'           # A/B switch
        if reserve if reserve is not None else (bin_thr is not None or not (dense and multi_label)):
            keys_tmp = torch.empty(capacity, device=dev, dtype=torch.int64)
            cand_tmp = torch.empty((capacity, 6), device=dev, dtype=torch.float32)
            f.update(keys_tmp=keys_tmp.data_ptr(), cand_tmp=cand_tmp.data_ptr())
        for i, lv in enumerate(levels):
            f[f"lv_logits{i}"] = lv.logits.data_ptr()
        if cm is not None:
            f["class_mask"] = cm.data_ptr()
        f["per_image_regions"] = int(PER_IMAGE_REGIONS)
        call("dmay_nms_filter_fused", s, **f)
        offs_host = img_offsets.tolist()     # the one sizing sync of the batch (N+1 values)
        total = offs_host[-1]
        maxc = max(offs_host[i + 1] - offs_host[i] for i in range(n))
        regions = PER_IMAGE_REGIONS and "keys_tmp" in f      # reserve mode: every image has capacity // n of the temporary buffers
        if total <= capacity and (not regions or maxc <= capacity // n):
            break
        # undersized guess: every candidate was counted, repeat once
        capacity = max(total + total // 8, (n * (maxc + maxc // 8)) if regions else 0) + 4096
    _FUSED_CAP[key] = max(total + total // 4 + 4096, (n * (maxc + maxc // 4) + 4096) if regions else 0, _FUSED_CAP.get(key, 0) // 2)
    if (not dense and multi_label and max_nms > 0 and FUSED_PRETHR_RATIO > 0 and not fused_pre
            and total > FUSED_PRETHR_RATIO * max_nms * n):
        _FUSED_PRETHR[key] = True     # from the next call on: only what the top-max_nms selection can keep is written
        _FUSED_CAP[key] = n * 2 * int(max_nms) + 4096
    return keys, cand, img_counts, img_offsets, offs_host, total


def nms_fused_static(levels, na: int, nc: int, conf_thres: float, iou_thres: float, capacity: int, *, classes=None,
                     agnostic=False, multi_label=False, max_det=300, max_nms=30000, max_wh=4096.0):
    """Sync-free form of the Detect-logits NMS chain for CUDA-graph capture: the candidate buffers have a FIXED `capacity`,
    nothing is read back by the host, every buffer size is a function of (capacity, N) only.
      filter (reserve + scan + gather) -> clamp offsets to capacity -> exact top-max_nms selection (a plain copy for images
      at or below max_nms) into a sentinel-filled key buffer -> stable sort of the whole buffer (sentinel keys sort last
      within the last image and are never read: the per-image counts bound the greedy kernel) -> greedy.
    Returns (out, counts, packed, img_offsets); img_offsets[N] is the TRUE candidate total: when it exceeds `capacity` the
    result of this call is invalid and the caller must repeat the step with larger buffers (GraphedDetector does)."""
    import ctypes
    dev = levels[0].logits.device
    n = levels[0].logits.shape[0]
    s = torch.cuda.current_stream(dev).cuda_stream
    multi_label = bool(multi_label) and nc > 1
    buf, out, out_counts = _det_buffers(n, max_det, dev)
    raw, rows = _level_meta_host(levels, na)
    meta_host = ctypes.create_string_buffer(raw, len(raw))
    ws_bytes = int(_lib.lib().dmay_nms_filter_fused_ws(ctypes.addressof(meta_host), len(levels), n))
    if ws_bytes < 0:
        raise DmayError(f"dmay_nms_filter_fused_ws failed: {ws_bytes}")
    cm = _class_mask(classes, nc, dev)
    capacity = int(capacity)
    img_counts = torch.empty(n, device=dev, dtype=torch.int32)
    img_offsets = torch.empty(n + 1, device=dev, dtype=torch.int64)
    ws = torch.zeros(ws_bytes // 8 + 1, device=dev, dtype=torch.int64)
    keys = torch.empty(capacity, device=dev, dtype=torch.int64)
    cand = torch.empty((capacity, 6), device=dev, dtype=torch.float32)
    keys_tmp = torch.empty(capacity, device=dev, dtype=torch.int64)
    cand_tmp = torch.empty((capacity, 6), device=dev, dtype=torch.float32)
    f = dict(lv_meta_host=ctypes.addressof(meta_host), ws=ws.data_ptr(), ws_bytes=ws.numel() * 8,
             img_counts=img_counts.data_ptr(), img_offsets=img_offsets.data_ptr(), keys=keys.data_ptr(),
             cand=cand.data_ptr(), N=n, nc=nc, levels=len(levels), multi_label=int(multi_label), capacity=capacity,
             conf_thres=float(conf_thres), keys_tmp=keys_tmp.data_ptr(), cand_tmp=cand_tmp.data_ptr(),
             row_pitch=levels[0].pitch, per_image_regions=int(PER_IMAGE_REGIONS))
    for i, lv in enumerate(levels):
        f[f"lv_logits{i}"] = lv.logits.data_ptr()
    if cm is not None:
        f["class_mask"] = cm.data_ptr()
    call("dmay_nms_filter_fused", s, **f)
    offs_cl = torch.empty(n + 1, device=dev, dtype=torch.int64)
    cnts_cl = torch.empty(n, device=dev, dtype=torch.int32)
    call("dmay_nms_clamp_offsets", s, img_offsets=img_offsets.data_ptr(), offsets_out=offs_cl.data_ptr(),
         counts_out=cnts_cl.data_ptr(), N=n, capacity=capacity)
    n_c = min(capacity, n * max_nms)
    keys_c = torch.full((n_c,), -1, device=dev, dtype=torch.int64)           # sentinel: sorts after every real key
    idx_c = torch.zeros(n_c, device=dev, dtype=torch.int32)
    counts_c = torch.empty(n, device=dev, dtype=torch.int32)
    offs_c = torch.empty(n + 1, device=dev, dtype=torch.int64)
    call("dmay_nms_topk_select", s, keys=keys.data_ptr(), img_counts=cnts_cl.data_ptr(), img_offsets=offs_cl.data_ptr(),
         keys_out=keys_c.data_ptr(), idx_out=idx_c.data_ptr(), counts_out=counts_c.data_ptr(), offsets_out=offs_c.data_ptr(),
         N=n, K=max_nms)
    _ko, idx = _sort_candidates(keys_c, n_c, n, dev, s, vals=idx_c)
    call("dmay_nms_greedy", s, cand=cand.data_ptr(), sorted_idx=idx.data_ptr(), img_counts=counts_c.data_ptr(),
         img_offsets=offs_c.data_ptr(), out=out.data_ptr(), out_counts=out_counts.data_ptr(), N=n, max_det=max_det,
         max_nms=max_nms, agnostic=int(bool(agnostic)), max_wh=float(max_wh), iou_thres=float(iou_thres))
    return out, out_counts, buf, img_offsets


def filter_candidates(pred, conf_thres, *, multi_label=False, classes=None, levels=None, na=0, nc=None, rows_kernel=False):
    """Candidate generation alone (tests, benchmarks): dense `pred` -> three-launch filter, Detect `levels` -> fused
    single-pass filter.  -> dict(keys, cand, img_counts, img_offsets (host list), total)."""
    if levels is not None:
        multi = bool(multi_label) and nc > 1
        keys, cand, img_counts, _, offs_host, total = _fused_candidates(levels, na, nc, conf_thres, multi, classes)
    else:
        if pred.dtype != torch.float32 or not pred.is_contiguous():
            pred = pred.float().contiguous()
        dev = pred.device
        n, rows, no = pred.shape
        nc = no - 5
        multi = bool(multi_label) and nc > 1
        s = torch.cuda.current_stream(dev).cuda_stream
        if rows_kernel and nc <= 96:
            keys, cand, img_counts, _, offs_host, total = _fused_candidates(_dense_level(pred), 1, nc, conf_thres, multi, classes,
                                                                            dense=True)
        else:
            keys, cand, img_counts, _, offs_host, total = _ordered_candidates(dict(levels=0, pred=pred.data_ptr()), n, rows, nc,
                                                                              conf_thres, multi, classes, dev, s)
    if total == 0:
        dev = img_counts.device
        keys = torch.empty(0, device=dev, dtype=torch.int64)
        cand = torch.empty((0, 6), device=dev, dtype=torch.float32)
    return dict(keys=keys, cand=cand, img_counts=img_counts, img_offsets=offs_host, total=total)


# --------------------------------------------------------------------------------------------
# 8f-2 HorBlock / GnConv pieces (C3HB)
# --------------------------------------------------------------------------------------------
def dwconv7(x: torch.Tensor, c_off: int, w49: torch.Tensor, bias: torch.Tensor, scale: float, seg_start, seg_out, c_out: int):
    """7x7 depth-wise conv over channels [c_off, c_off + Cd) of the NHWC tensor x; segment i of the result lands at channel
    seg_out[i] of a [N, c_out, H, W] tensor (zero-initialised when the segments leave gaps)."""
    x = as_act(x)
    n, _, h, w = x.shape
    cd = w49.shape[1]
    dense = c_out == cd
    y = empty_nhwc(n, c_out, h, w, x.device) if dense else torch.zeros((n, h, w, c_out), device=x.device,
                                                                       dtype=torch.bfloat16).permute(0, 3, 1, 2)
    f = dict(x=x.data_ptr() + 2 * c_off, w=w49.data_ptr(), bias=bias.data_ptr(), y=y.data_ptr(), N=n, H=h, W=w, Cd=cd,
             ldx=ld_of(x), ldy=ld_of(y), scale=float(scale), n_seg=len(seg_start))
    for i, (a, b) in enumerate(zip(seg_start, seg_out)):
        f[f"seg_start{i}"] = int(a)
        f[f"seg_out{i}"] = int(b)
    call("dmay_dwconv7", _stream(x), **f)
    return y


def mul_channels(a: torch.Tensor, b: torch.Tensor, d: int, c_alloc: int):
    """[N, d] product of the first d channels of two NHWC tensors, in a zero-padded [N, c_alloc, H, W] tensor."""
    n, _, h, w = a.shape
    y = torch.zeros((n, h, w, c_alloc), device=a.device, dtype=torch.bfloat16).permute(0, 3, 1, 2)
    call("dmay_mul_channels", _stream(a), a=a.data_ptr(), b=b.data_ptr(), y=y.data_ptr(), npix=n * h * w, d=d, lda=ld_of(a),
         ldb=ld_of(b), ldy=c_alloc)
    return y


def axpy_channels(x: torch.Tensor, g: torch.Tensor, gamma: torch.Tensor, out=None):
    """x + gamma[c] * g."""
    x, g = as_act(x), as_act(g)
    n, c, h, w = x.shape
    if out is None:
        out = empty_nhwc(n, c, h, w, x.device)
    call("dmay_axpy_channels", _stream(x), x=x.data_ptr(), g=g.data_ptr(), gamma=gamma.data_ptr(), y=out.data_ptr(),
         npix=n * h * w, C=c, ldx=ld_of(x), ldg=ld_of(g), ldy=ld_of(out))
    return out


def device_copy(dst: torch.Tensor, src: torch.Tensor):
    call("dmay_copy", _stream(src), src=src.data_ptr(), dst=dst.data_ptr(), bytes=src.numel() * src.element_size())


# --------------------------------------------------------------------------------------------
# 8f-4 TDetect eval tail: DFL expectation + dist2bbox + sigmoid in one kernel
# --------------------------------------------------------------------------------------------
def dfl_decode(boxes, clss, nc: int, reg_max: int, strides) -> torch.Tensor:
    """Per-level fp32 NHWC head logits (`boxes[i]` [N, 4*reg_max, ny, nx], `clss[i]` [N, nc, ny, nx], channels_last strides)
    -> y [N, 4 + nc, A] exactly as TDetect.forward's eval branch builds it (models/detect_t.py:46-58)."""
    n = boxes[0].shape[0]
    A = sum(b.shape[2] * b.shape[3] for b in boxes)
    y = torch.empty((n, 4 + nc, A), device=boxes[0].device, dtype=torch.float32)
    a0 = 0
    for b, c, st in zip(boxes, clss, strides):
        if b.dtype != torch.float32 or c.dtype != torch.float32 or not is_nhwc(b) or not is_nhwc(c):
            raise DmayError("dfl_decode: fp32 NHWC head logits expected")
        ny, nx = b.shape[2], b.shape[3]
        call("dmay_dfl_decode", _stream(b), box=b.data_ptr(), cls=c.data_ptr(), y=y.data_ptr(), N=n, ny=ny, nx=nx, nc=nc,
             reg_max=reg_max, ld_box=ld_of(b), ld_cls=ld_of(c), a0=a0, A=A, stride=float(st))
        a0 += ny * nx
    return y
