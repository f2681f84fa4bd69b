"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — CPU fp32 restatement of the reference blocks.

Plain functional PyTorch on CPU tensors: every function takes the block's weights explicitly (a
state_dict `sd` + key prefix, Appendix B key names) and cites the reference lines it restates.
`forward_model` interprets a model YAML with these functions, i.e. it is an independent
re-derivation of `Model._forward_once` that shares no code with `dma_yolo_b200`.
"""
from __future__ import annotations

import contextlib
import math

import torch
import torch.nn.functional as F

# bf16-storage emulation.  The reference in fp32 is the default oracle.  Under `bf16_storage()` every tensor that the
# B200 path STORES as bf16 between two fused kernels is rounded to bf16 here as well (and only there: inside a fused
# kernel — conv accumulator -> BN -> SiLU -> residual / gate — everything stays fp32, exactly one rounding at the
# end).  This is the oracle for the "within 1e-2 in bf16" tolerance: it removes the storage rounding, which
# compounds over 25 layers of an untrained net, from the comparison and leaves only arithmetic differences
# (accumulation order, tanh.approx SiLU).
_BF16 = False


@contextlib.contextmanager
def bf16_storage(on=True):
    global _BF16
    old, _BF16 = _BF16, on
    try:
        yield
    finally:
        _BF16 = old


def q(t):
    return t.bfloat16().float() if _BF16 else t


def _bn(x, sd, p, eps):
    """nn.BatchNorm2d in eval mode with the module's own eps (utils/torch_utils.py:167 sets 1e-3 inside a Model)."""
    return F.batch_norm(x, sd[p + 'running_mean'], sd[p + 'running_var'], sd[p + 'weight'], sd[p + 'bias'], False, 0.0, eps)


def silu(x):
    return x * torch.sigmoid(x)


def conv_bn_act(x, sd, p, k=1, s=1, pad=None, act=True, eps=1e-3, raw=False):
    """Conv.forward — models/common.py:50-77: act(bn(conv(x))), conv bias=False, pad=k//2.
    After Model.fuse() (models/yolo.py:315-323) the module has a biased conv and no bn."""
    pad = k // 2 if pad is None else pad
    if p + 'bn.weight' in sd:
        y = _bn(F.conv2d(x, sd[p + 'conv.weight'], None, s, pad), sd, p + 'bn.', eps)
    else:
        y = F.conv2d(x, sd[p + 'conv.weight'], sd.get(p + 'conv.bias'), s, pad)
    y = silu(y) if act else y
    return y if raw else q(y)


def bottleneck(x, sd, p, shortcut=True, eps=1e-3):
    """Bottleneck.forward — models/common.py:119-137 (e=1.0 inside C3): x + cv2(cv1(x))."""
    c1, c2 = sd[p + 'cv1.conv.weight'].shape[1], sd[p + 'cv2.conv.weight'].shape[0]
    add = shortcut and c1 == c2
    y = conv_bn_act(conv_bn_act(x, sd, p + 'cv1.', 1, eps=eps), sd, p + 'cv2.', 3, eps=eps, raw=add)
    return q(x + y) if add else y      # the residual add lives in cv2's epilogue: one rounding after the sum


def c3(x, sd, p, n=1, shortcut=True, eps=1e-3):
    """C3.forward — models/common.py:159-182: cv3(cat(m(cv1 x), cv2 x))."""
    y = conv_bn_act(x, sd, p + 'cv1.', 1, eps=eps)
    for i in range(n):
        y = bottleneck(y, sd, f'{p}m.{i}.', shortcut, eps)
    return conv_bn_act(torch.cat((y, conv_bn_act(x, sd, p + 'cv2.', 1, eps=eps)), 1), sd, p + 'cv3.', 1, eps=eps)


def coordatt(x, sd, p, eps=1e-3):
    """CoorAttention.forward — models/common.py:1183-1207 (restatement SURVEY.md App. D-4)."""
    n, c, h, w = x.shape
    ph = x.mean(dim=3, keepdim=True)                       # [n,c,h,1]
    pw = x.mean(dim=2, keepdim=True).permute(0, 1, 3, 2)   # [n,c,w,1]
    y = F.conv2d(torch.cat([ph, pw], 2), sd[p + 'conv1.weight'], sd[p + 'conv1.bias'])
    y = F.hardswish(_bn(y, sd, p + 'bn1.', eps))
    yh, yw = torch.split(y, [h, w], dim=2)
    a_h = torch.sigmoid(F.conv2d(yh, sd[p + 'conv_h.weight'], sd[p + 'conv_h.bias']))                      # [n,c,h,1]
    a_w = torch.sigmoid(F.conv2d(yw.permute(0, 1, 3, 2), sd[p + 'conv_w.weight'], sd[p + 'conv_w.bias']))  # [n,c,1,w]
    return q(x * a_w * a_h)


def space_to_depth(x):
    """space_to_depth.forward — models/common.py:1457-1458: out[n,(dy+2dx)C+c,h,w] = in[n,c,2h+dy,2w+dx]."""
    return torch.cat([x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]], 1)


def nearest_index(dst: int, in_size: int, out_size: int) -> int:
    """ATen nearest source index: min(floor(dst * float(in)/float(out)), in-1) in fp32."""
    import numpy as np
    scale = np.float32(in_size) / np.float32(out_size)
    return min(int(np.floor(np.float32(dst) * scale)), in_size - 1)


def scconv_gate(x, k3o, k2o):
    """k3(x) * sigmoid(x + nearest_up(k2)) — models/common.py:1310-1314, explicit index form (App. D-7)."""
    H, W = x.shape[2:]
    Hk, Wk = k2o.shape[2:]
    hi = torch.tensor([nearest_index(h, Hk, H) for h in range(H)])
    wi = torch.tensor([nearest_index(w, Wk, W) for w in range(W)])
    up = k2o[:, :, hi][:, :, :, wi]
    return k3o * torch.sigmoid(x + up)


def scconv(x, sd, p, stride, pooling_r=4, eps=1e-3):
    """SCConv.forward — models/common.py:1279-1316 (no activation anywhere inside)."""
    k2o = q(_bn(F.conv2d(q(F.avg_pool2d(x, pooling_r, pooling_r)), sd[p + 'k2.1.weight'], None, 1, 1), sd, p + 'k2.2.', eps))
    k3o = _bn(F.conv2d(x, sd[p + 'k3.0.weight'], None, 1, 1), sd, p + 'k3.1.', eps)   # stays fp32: the gate is k3's epilogue
    g = q(scconv_gate(x, k3o, k2o))
    return q(_bn(F.conv2d(g, sd[p + 'k4.0.weight'], None, stride, 1), sd, p + 'k4.1.', eps))


def adconcat(xs, w, epsilon=1e-4):
    """AdConcat2/3.forward — models/common.py:1003-1008,1021-1026."""
    weight = w / (torch.sum(w, dim=0) + epsilon)
    return torch.cat([q(weight[i] * x) for i, x in enumerate(xs)], 1)


def adapt_add2(xs, w, epsilon=1e-4):
    """Adapt_Add2.forward — models/common.py:1040-1045."""
    weight = w / (torch.sum(w, dim=0) + epsilon)
    return q(silu(weight[0] * xs[0] + weight[1] * xs[1]))


def adapt_add3(xs, sd, p, epsilon=1e-4):
    """Adapt_Add3.forward — models/common.py:1056-1061 (shared biased 1x1 conv on the first two inputs)."""
    w = sd[p + 'w']
    weight = w / (torch.sum(w, dim=0) + epsilon)
    cv = lambda t: F.conv2d(t, sd[p + 'conv.weight'], sd[p + 'conv.bias'])
    return q(silu(weight[0] * cv(xs[0]) + weight[1] * cv(xs[1]) + weight[2] * xs[2]))


def maxpool_cascade(x, k=5):
    """x -> (mp_k x, mp_k mp_k x, mp_k mp_k mp_k x), stride 1, -inf padding — models/common.py:1272-1274."""
    y1 = F.max_pool2d(x, k, 1, k // 2)
    y2 = F.max_pool2d(y1, k, 1, k // 2)
    return y1, y2, F.max_pool2d(y2, k, 1, k // 2)


def sppf(x, sd, p, k=5, eps=1e-3):
    """SPPF.forward — models/common.py:252-258."""
    x1 = conv_bn_act(x, sd, p + 'cv1.', 1, eps=eps)
    return conv_bn_act(torch.cat((x1,) + maxpool_cascade(x1, k), 1), sd, p + 'cv2.', 1, eps=eps)


def spp(x, sd, p, ks=(5, 9, 13), eps=1e-3):
    """SPP.forward — models/common.py:221-227."""
    x1 = conv_bn_act(x, sd, p + 'cv1.', 1, eps=eps)
    return conv_bn_act(torch.cat([x1] + [F.max_pool2d(x1, k, 1, k // 2) for k in ks], 1), sd, p + 'cv2.', 1, eps=eps)


def sppfcspc(x, sd, p, k=5, eps=1e-3):
    """SPPFCSPC.forward — models/common.py:1270-1276."""
    cba = lambda t, name, kk: conv_bn_act(t, sd, f'{p}{name}.', kk, eps=eps)
    x1 = cba(cba(cba(x, 'cv1', 1), 'cv3', 3), 'cv4', 1)
    y1 = cba(cba(torch.cat((x1,) + maxpool_cascade(x1, k), 1), 'cv5', 1), 'cv6', 3)
    return cba(torch.cat((y1, cba(x, 'cv2', 1)), 1), 'cv7', 1)


def sppcspc(x, sd, p, ks=(5, 9, 13), eps=1e-3):
    """SPPCSPC.forward — models/common.py:1250-1255."""
    cba = lambda t, name, kk: conv_bn_act(t, sd, f'{p}{name}.', kk, eps=eps)
    x1 = cba(cba(cba(x, 'cv1', 1), 'cv3', 3), 'cv4', 1)
    y1 = cba(cba(torch.cat([x1] + [F.max_pool2d(x1, k, 1, k // 2) for k in ks], 1), 'cv5', 1), 'cv6', 3)
    return cba(torch.cat((y1, cba(x, 'cv2', 1)), 1), 'cv7', 1)


def detect(xs, sd, p, anchors_grid, strides, nc):
    """Detect.forward (eval) — models/yolo.py:68-103.  anchors_grid = the `anchors` buffer (grid units).
    Returns (pred [bs, rows, no], [x_i (bs,na,ny,nx,no) raw logits])."""
    no = nc + 5
    na = anchors_grid.shape[1]
    z, raw = [], []
    for i, x in enumerate(xs):
        t = F.conv2d(x, sd[f'{p}m.{i}.weight'], sd[f'{p}m.{i}.bias'])
        bs, _, ny, nx = t.shape
        t = t.view(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
        raw.append(t)
        yv, xv = torch.meshgrid(torch.arange(ny), torch.arange(nx), indexing='ij')
        grid = torch.stack((xv, yv), 2).view(1, 1, ny, nx, 2).float()
        ag = (anchors_grid[i].float() * strides[i]).view(1, na, 1, 1, 2)
        y = t.sigmoid()
        xy = (y[..., 0:2] * 2 - 0.5 + grid) * strides[i]
        wh = (y[..., 2:4] * 2) ** 2 * ag
        z.append(torch.cat((xy, wh, y[..., 4:]), -1).view(bs, -1, no))
    return torch.cat(z, 1), raw


def tdetect(xs, sd, p, nc, strides, reg_max=16, eps=1e-3):
    """TDetect.forward (eval) — models/detect_t.py:23-58 with DFL 92-102, make_anchors 66-79, dist2bbox 81-90.
    Per level: box = 1x1(Conv3(Conv3(x))) [4*reg_max], cls = 1x1(Conv3(Conv3(x))) [nc]; over all levels the softmax
    expectation of the `reg_max` bins gives (l, t, r, b) distances from the cell centre (+0.5) in grid units.
    Returns (y [b, 4 + nc, A], box [b, 4*reg_max, A], cls [b, nc, A])."""
    feats = []
    for i, x in enumerate(xs):
        outs = []
        for br in ('cv2', 'cv3'):
            t = conv_bn_act(x, sd, f'{p}{br}.{i}.0.', 3, 1, eps=eps)
            t = conv_bn_act(t, sd, f'{p}{br}.{i}.1.', 3, 1, eps=eps)
            outs.append(F.conv2d(t, sd[f'{p}{br}.{i}.2.weight'], sd[f'{p}{br}.{i}.2.bias']))
        feats.append(torch.cat(outs, 1))
    b = feats[0].shape[0]
    no = 4 * reg_max + nc
    allf = torch.cat([f.reshape(b, no, -1) for f in feats], 2)
    box, cls = allf[:, :4 * reg_max], allf[:, 4 * reg_max:]
    pts, strs = [], []
    for f, st in zip(feats, strides):
        h, w = f.shape[-2:]
        gy, gx = torch.meshgrid(torch.arange(h).float() + 0.5, torch.arange(w).float() + 0.5, indexing='ij')
        pts.append(torch.stack((gx, gy), -1).view(-1, 2))
        strs.append(torch.full((h * w,), float(st)))
    anchor = torch.cat(pts).t().unsqueeze(0)                  # [1, 2, A]
    stride = torch.cat(strs).view(1, 1, -1)
    prob = box.view(b, 4, reg_max, -1).softmax(2)
    dist = (prob * torch.arange(reg_max).float().view(1, 1, reg_max, 1)).sum(2)     # [b, 4, A]
    lt, rb = dist[:, :2], dist[:, 2:]
    x1y1, x2y2 = anchor - lt, anchor + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * stride
    return torch.cat((dbox, cls.sigmoid()), 1), box, cls


def make_divisible(x, d):
    return math.ceil(x / d) * d


def forward_model(cfg: dict, sd: dict, x: torch.Tensor, strides, eps=1e-3, teacher=None):
    """Model._forward_once (eval) — models/yolo.py:211-239 + the channel/arg bookkeeping of parse_model
    (models/yolo.py:353-478), as a functional interpreter over the hot-path module set.
    `sd` = state_dict of the model (keys `model.{i}....`).  Returns (pred, raw_list, per-layer outputs).
    `teacher` (a list of per-layer tensors): every layer reads ITS inputs from that list instead of from this run's
    own outputs (teacher forcing — used to measure one layer's sensitivity to storage precision on given inputs)."""
    gd, nc = cfg['depth_multiple'], cfg['nc']
    ys, outs = [], []
    x = q(x)
    for i, (f, n, m, args) in enumerate(cfg['backbone'] + cfg['head']):
        p = f'model.{i}.'
        if f != -1:
            xin = ys[f] if isinstance(f, int) else [x if j == -1 else ys[j] for j in f]
        else:
            xin = x
        n = max(round(n * gd), 1) if n > 1 else n
        a = [None if v == 'None' else v for v in args]
        if m == 'Conv':
            k = a[1] if len(a) > 1 else 1
            s = a[2] if len(a) > 2 else 1
            pad = a[3] if len(a) > 3 else None
            if n > 1:
                for r in range(n):
                    xin = conv_bn_act(xin, sd, f'{p}{r}.', k, s, pad, eps=eps)
                x = xin
            else:
                x = conv_bn_act(xin, sd, p, k, s, pad, eps=eps)
        elif m == 'C3':
            x = c3(xin, sd, p, n, a[1] if len(a) > 1 else True, eps)
        elif m == 'C3STR':
            x = c3str(xin, sd, p, n, eps)
        elif m == 'C3CA':
            x = c3ca(xin, sd, p, n, a[1] if len(a) > 1 else True, eps)
        elif m == 'C3HB':
            x = c3hb(xin, sd, p, n, eps)
        elif m == 'SCConv':
            x = scconv(xin, sd, p, a[1], eps=eps)
        elif m in ('CA', 'CoorAttention'):
            x = coordatt(xin, sd, p, eps)
        elif m == 'SPPFCSPC':
            x = sppfcspc(xin, sd, p, eps=eps)
        elif m == 'SPPCSPC':
            x = sppcspc(xin, sd, p, eps=eps)
        elif m == 'SPPF':
            x = sppf(xin, sd, p, a[1] if len(a) > 1 else 5, eps)
        elif m == 'SPP':
            x = spp(xin, sd, p, tuple(a[1]) if len(a) > 1 else (5, 9, 13), eps)
        elif m == 'nn.Upsample':
            x = F.interpolate(xin, scale_factor=a[1], mode=a[2])
        elif m in ('AdConcat2', 'AdConcat3'):
            x = adconcat(xin, sd[p + 'w'])
        elif m == 'Adapt_Add2':
            x = adapt_add2(xin, sd[p + 'w'])
        elif m == 'Concat':
            x = torch.cat(xin, 1)
        elif m in ('space_to_depth', 'SM'):
            x = space_to_depth(xin)
        elif m == 'Focus':
            x = conv_bn_act(space_to_depth(xin), sd, p + 'conv.', a[1] if len(a) > 1 else 1, eps=eps)
        elif m == 'Detect':
            pred, raw = detect(xin, sd, p, sd[p + 'anchors'], strides, nc)
            outs.append(pred)
            return pred, raw, outs
        else:
            raise NotImplementedError(f'oracle: layer {i} module {m}')
        outs.append(x)
        if teacher is not None:
            x = teacher[i]
        ys.append(x)
    return x, None, outs


# ---- 8f-1: Swin transformer layer / C3STR (models/common.py:448-654, 191-196) -------------------------------------
def swin_shift_labels(Rp, Sp, ws, ss):
    """Region labels of SwinTransformerLayer.create_mask (models/common.py:567-591) on the padded (rows, cols) grid.
    The reference's first row selector is the TUPLE (0, -ws): it labels the two rows 0 and Rp-ws, not a range; later
    assignments overwrite earlier ones."""
    lab = torch.zeros(Rp, Sp)
    row_sets = ([0, Rp - ws], list(range(Rp - ws, Rp - ss)), list(range(Rp - ss, Rp)))
    col_sets = (list(range(0, Sp - ws)), list(range(Sp - ws, Sp - ss)), list(range(Sp - ss, Sp)))
    cnt = 0
    for rows in row_sets:
        for cols in col_sets:
            for r in rows:
                for c in cols:
                    lab[r, c] = cnt
            cnt += 1
    return lab


def swin_layer(x, sd, p, heads, ws=8, shift=0, eps=1e-5):
    """SwinTransformerLayer.forward (models/common.py:593-634) with explicit index arithmetic.  x is NCHW; the
    reference reads it as (b, c, w, h) and works on the transposed map (b, h, w, c)."""
    b, c, D2, D3 = x.shape
    hd = c // heads
    t = x.permute(0, 3, 2, 1)                                    # rows R = D3, cols S = D2
    R, S = D3, D2
    Rp, Sp = -(-R // ws) * ws, -(-S // ws) * ws
    y = q(F.layer_norm(t, (c,), sd[p + 'norm1.weight'], sd[p + 'norm1.bias'], eps))
    yp = torch.zeros(b, Rp, Sp, c)
    yp[:, :R, :S] = y                                            # zero padding AFTER the norm
    qkv = q(yp @ sd[p + 'attn.qkv.weight'].t())                  # no bias: padded tokens give q = k = v = 0
    table, index = sd[p + 'attn.relative_position_bias_table'], sd[p + 'attn.relative_position_index']
    n_tok = ws * ws
    bias = table[index.reshape(-1).long()].reshape(n_tok, n_tok, heads).permute(2, 0, 1)   # [heads, N, N]
    lab = swin_shift_labels(Rp, Sp, ws, shift) if shift > 0 else None
    out = torch.zeros(b, Rp, Sp, c)
    scale = hd ** -0.5
    ri = torch.arange(ws).repeat_interleave(ws)
    si = torch.arange(ws).repeat(ws)
    for wr in range(Rp // ws):
        for wc in range(Sp // ws):
            rows = (wr * ws + ri + shift) % Rp                   # roll(-shift): shifted[i] = x[(i + shift) % n]
            cols = (wc * ws + si + shift) % Sp
            tok = qkv[:, rows, cols]                             # [b, N, 3c]
            qq, kk, vv = tok.reshape(b, n_tok, 3, heads, hd).permute(2, 0, 3, 1, 4)
            a = (qq * scale) @ kk.transpose(-2, -1) + bias[None]
            if lab is not None:
                l = lab[wr * ws + ri, wc * ws + si]              # labels are defined on the SHIFTED grid
                a = a + torch.where(l[None, :] != l[:, None], torch.tensor(-100.0), torch.tensor(0.0))[None, None]
            o = torch.softmax(a, -1) @ vv                        # [b, heads, N, hd]
            out[:, rows, cols] = o.permute(0, 2, 1, 3).reshape(b, n_tok, c)
    att = q(out[:, :R, :S])
    x1 = q(t + att @ sd[p + 'attn.proj.weight'].t() + sd[p + 'attn.proj.bias'])
    y2 = q(F.layer_norm(x1, (c,), sd[p + 'norm2.weight'], sd[p + 'norm2.bias'], eps))
    hdn = q(F.gelu(y2 @ sd[p + 'mlp.fc1.weight'].t() + sd[p + 'mlp.fc1.bias']))
    x2 = q(x1 + hdn @ sd[p + 'mlp.fc2.weight'].t() + sd[p + 'mlp.fc2.bias'])
    return x2.permute(0, 3, 2, 1).contiguous()


def c3str(x, sd, p, n=1, eps=1e-3, ws=8):
    """C3STR.forward — models/common.py:191-196 on top of C3 (159-182): cv3(cat(swin(cv1 x), cv2 x))."""
    y = conv_bn_act(x, sd, p + 'cv1.', 1, eps=eps)
    heads = y.shape[1] // 32
    for i in range(n):
        y = swin_layer(y, sd, f'{p}m.tr.{i}.', heads, ws, 0 if i % 2 == 0 else ws // 2)
    return conv_bn_act(torch.cat((y, conv_bn_act(x, sd, p + 'cv2.', 1, eps=eps)), 1), sd, p + 'cv3.', 1, eps=eps)


def ca_bottleneck(x, sd, p, shortcut=True, eps=1e-3):
    """CABottleneck.forward — models/common.py:1209-1227 (e=1.0 inside C3CA): x + ca(cv2(cv1(x)))."""
    y = coordatt(conv_bn_act(conv_bn_act(x, sd, p + 'cv1.', 1, eps=eps), sd, p + 'cv2.', 3, eps=eps), sd, p + 'ca.', eps)
    c1, c2 = sd[p + 'cv1.conv.weight'].shape[1], sd[p + 'cv2.conv.weight'].shape[0]
    return q(x + y) if (shortcut and c1 == c2) else y


def c3ca(x, sd, p, n=1, shortcut=True, eps=1e-3):
    """C3CA.forward — models/common.py:1229-1235 on top of C3 (159-182)."""
    y = conv_bn_act(x, sd, p + 'cv1.', 1, eps=eps)
    for i in range(n):
        y = ca_bottleneck(y, sd, f'{p}m.{i}.', shortcut, eps)
    return conv_bn_act(torch.cat((y, conv_bn_act(x, sd, p + 'cv2.', 1, eps=eps)), 1), sd, p + 'cv3.', 1, eps=eps)


# ---- 8f-2: HorBlock / GnConv / C3HB (models/common.py:1318-1426) --------------------------------------------------
def gnconv(x, sd, p, order=5, scale=1.0, eps=1e-3):
    """GnConv.forward — models/common.py:1336-1346: recursive gating over channel groups c/16, c/8, ..., c."""
    c = sd[p + 'proj_in.weight'].shape[1]
    dims = [c // 2 ** i for i in range(order)][::-1]
    fused = q(F.conv2d(x, sd[p + 'proj_in.weight'], sd[p + 'proj_in.bias']))
    pwa, abc = fused[:, :dims[0]], fused[:, dims[0]:]
    dw = q(F.conv2d(abc, sd[p + 'dwconv.weight'], sd[p + 'dwconv.bias'], padding=3, groups=sum(dims)) * scale)
    offs = [sum(dims[:i]) for i in range(order + 1)]
    y = q(pwa * dw[:, offs[0]:offs[1]])
    for i in range(order - 1):
        y = q(F.conv2d(y, sd[f'{p}pws.{i}.weight'], sd[f'{p}pws.{i}.bias']) * dw[:, offs[i + 1]:offs[i + 2]])
    return conv_bn_act(y, sd, p + 'proj_out.', 1, eps=eps)


def horblock(x, sd, p, eps=1e-3):
    """HorBlock.forward — models/common.py:1367-1383 (channels_first LayerNorm 1397-1409 written out)."""
    c = x.shape[1]
    u = x.mean(1, keepdim=True)
    s = ((x - u) ** 2).mean(1, keepdim=True)
    y = q(sd[p + 'norm1.weight'][:, None, None] * ((x - u) / torch.sqrt(s + 1e-6)) + sd[p + 'norm1.bias'][:, None, None])
    g = gnconv(y, sd, p + 'gnconv.', eps=eps)
    x1 = q(x + sd[p + 'gamma1'].view(c, 1, 1) * g)
    t = x1.permute(0, 2, 3, 1)
    y2 = q(F.layer_norm(t, (c,), sd[p + 'norm2.weight'], sd[p + 'norm2.bias'], 1e-6))
    h = q(F.gelu(y2 @ sd[p + 'pwconv1.weight'].t() + sd[p + 'pwconv1.bias']))
    o = sd[p + 'gamma2'] * (h @ sd[p + 'pwconv2.weight'].t() + sd[p + 'pwconv2.bias'])
    return q(x1 + o.permute(0, 3, 1, 2))


def c3hb(x, sd, p, n=1, eps=1e-3):
    """C3HB.forward — models/common.py:1412-1426."""
    y = conv_bn_act(x, sd, p + 'cv1.', 1, eps=eps)
    for i in range(n):
        y = horblock(y, sd, f'{p}m.{i}.', eps)
    return conv_bn_act(torch.cat((y, conv_bn_act(x, sd, p + 'cv2.', 1, eps=eps)), 1), sd, p + 'cv3.', 1, eps=eps)


# ---- 8f-4: test-time augmentation (models/yolo.py:194-209, 241-275; scale_img utils/torch_utils.py:264-274) -----------
def scale_img(img, ratio=1.0, gs=32):
    """Bilinear resize by `ratio`, padded (value 0.447) to the next multiple of the grid size."""
    if ratio == 1.0:
        return img
    h, w = img.shape[2:]
    s = (int(h * ratio), int(w * ratio))
    img = F.interpolate(img, size=s, mode='bilinear', align_corners=False)
    hp, wp = (math.ceil(v * ratio / gs) * gs for v in (h, w))
    return F.pad(img, [0, wp - s[1], 0, hp - s[0]], value=0.447)


def forward_augment(cfg: dict, sd: dict, x: torch.Tensor, strides, nl: int):
    """Model._forward_augment: six passes (scales 1 / 0.83 / 0.67, each plain and left-right flipped), predictions de-scaled
    and de-flipped, the tail of the first pass (largest-stride level) and the head of the last pass (smallest-stride level)
    clipped, everything concatenated along the row axis."""
    img_h, img_w = x.shape[-2:]
    ys = []
    for si, fi in zip([1, 1, 0.83, 0.83, 0.67, 0.67], [None, 3, None, 3, None, 3]):
        xi = scale_img(x.flip(fi) if fi else x, si, gs=int(max(strides)))
        p = forward_model(cfg, sd, xi, strides)[0].clone()
        p[..., :4] /= si
        if fi == 3:
            p[..., 0] = img_w - p[..., 0]
        ys.append(p)
    g = sum(4 ** k for k in range(nl))
    ys[0] = ys[0][:, :-(ys[0].shape[1] // g)]
    ys[-1] = ys[-1][:, (ys[-1].shape[1] // g) * 4 ** (nl - 1):]
    return torch.cat(ys, 1)
