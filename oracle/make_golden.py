"""Generate tests/golden/*.npz by EXECUTING the unmodified reference (build container only).

    python -m oracle.make_golden

Each block fixture holds the bf16-rounded input(s), the block's state_dict (bf16-rounded floats) and the
reference's fp32 output on CPU.  Model fixtures hold the seed recipe, a digest of the resulting
state_dict and the reference outputs, so the GPU box (where /root/reference does not exist) can rebuild
the same weights from the seed and compare.  NMS fixtures hold inputs and the reference
`non_max_suppression` outputs (called one image at a time).
"""
from __future__ import annotations

import hashlib
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / 'tests' / 'golden'


def bf16_round_(t: torch.Tensor):
    t.data = t.data.bfloat16().float()


def randomize_bn(mod: nn.Module, g: torch.Generator):
    for m in mod.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.eps = 1e-3  # as inside a Model (utils/torch_utils.py:167)
            m.weight.data = torch.rand(m.weight.shape, generator=g) + 0.5
            m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.2
            m.running_mean.data = torch.randn(m.running_mean.shape, generator=g) * 0.2
            m.running_var.data = torch.rand(m.running_var.shape, generator=g) + 0.5


def round_module(mod: nn.Module):
    for p in mod.parameters():
        bf16_round_(p)
    for n, b in mod.named_buffers():
        if b.is_floating_point():
            bf16_round_(b)


def state_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        v = sd[k]
        if torch.is_tensor(v):
            h.update(k.encode())
            h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build_calibrated(Model, cfg, seed=0, nc=None, calib_hw=(320, 320), calib_bs=4):
    """seed -> build -> BN momentum 1 -> one train-mode forward -> eval; then
    round every float parameter/buffer to bf16 so both sides hold identical weights."""
    torch.manual_seed(seed)
    m = Model(cfg, nc=nc) if nc else Model(cfg)
    for mod in m.modules():
        if isinstance(mod, nn.BatchNorm2d):
            mod.momentum = 1.0
    m.train()
    with torch.no_grad():
        m(torch.rand(calib_bs, 3, *calib_hw, generator=torch.Generator().manual_seed(seed + 1)))
    m.eval()
    for mod in m.modules():
        if isinstance(mod, nn.BatchNorm2d):
            mod.momentum = 0.03
    round_module(m)
    return m


def save(name, **arrs):
    GOLD.mkdir(parents=True, exist_ok=True)
    out = {}
    for k, v in arrs.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = v
    np.savez_compressed(GOLD / f'{name}.npz', **out)
    print(f'{name}: {os.path.getsize(GOLD / (name + ".npz")) / 1024:.1f} KiB')


def block_fixture(name, mod, inputs, seed):
    g = torch.Generator().manual_seed(seed)
    randomize_bn(mod, g)
    round_module(mod)
    mod.eval()
    with torch.no_grad():
        out = mod([t.clone() for t in inputs] if len(inputs) > 1 else inputs[0].clone())
    arrs = {f'in{i}': t for i, t in enumerate(inputs)}
    arrs['out'] = out
    for k, v in mod.state_dict().items():
        arrs['sd/' + k] = v
    save(name, **arrs)


def rnd(g, *shape, scale=1.0):
    return (torch.randn(*shape, generator=g) * scale).bfloat16().float()


def main(only=None):
    sys.path.insert(0, str(ROOT))
    from oracle import refshim
    R = refshim.load()
    C, Y = R.common, R.yolo
    g = torch.Generator().manual_seed(1234)
    torch.manual_seed(1234)

    # ---- a1/a2 convs, bottleneck, C3 ----
    block_fixture('conv_k3s1', C.Conv(16, 32, 3, 1), [rnd(g, 2, 16, 12, 10)], 1)
    block_fixture('conv_k3s2_odd', C.Conv(16, 32, 3, 2), [rnd(g, 2, 16, 11, 9)], 2)
    block_fixture('conv_k1', C.Conv(32, 16, 1, 1), [rnd(g, 2, 32, 6, 5)], 3)
    block_fixture('conv_stem_k6s2p2', C.Conv(3, 32, 6, 2, 2), [torch.rand(2, 3, 32, 24, generator=g).bfloat16().float()], 4)
    block_fixture('conv_c64_k3', C.Conv(64, 64, 3, 1), [rnd(g, 1, 64, 9, 7)], 5)
    block_fixture('bottleneck', C.Bottleneck(32, 32, True, e=1.0), [rnd(g, 2, 32, 8, 8)], 6)
    block_fixture('c3_n2', C.C3(32, 32, 2), [rnd(g, 2, 32, 8, 8)], 7)
    block_fixture('c3_n1_noshortcut', C.C3(64, 32, 1, False), [rnd(g, 1, 64, 6, 10)], 8)
    # ---- a3 CoordAtt (H != W) ----
    block_fixture('coordatt_7x5', C.CoorAttention(64, 64), [rnd(g, 2, 64, 7, 5)], 9)
    block_fixture('coordatt_20x20', C.CoorAttention(256, 256), [rnd(g, 1, 256, 20, 20)], 10)
    # ---- a4 SPD ----
    block_fixture('spd', C.space_to_depth(), [rnd(g, 2, 16, 8, 6)], 11)
    # ---- a5 SCConv (non-divisible sizes: 38x38 -> k2 9x9; 16x12 -> 4x3; 19x23) ----
    block_fixture('scconv_38', C.SCConv(16, 32, 2), [rnd(g, 1, 16, 38, 38)], 12)
    block_fixture('scconv_16x12', C.SCConv(32, 64, 2), [rnd(g, 2, 32, 16, 12)], 13)
    block_fixture('scconv_19x23_s1', C.SCConv(16, 16, 1), [rnd(g, 1, 16, 19, 23)], 14)
    # ---- a6 BiFPN fusion ----
    ad2 = C.AdConcat2(); ad2.w.data = torch.tensor([0.7, 1.6])
    block_fixture('adconcat2', ad2, [rnd(g, 2, 16, 6, 5), rnd(g, 2, 32, 6, 5)], 15)
    ad3 = C.AdConcat3(); ad3.w.data = torch.tensor([1.2, 0.4, 0.9])
    block_fixture('adconcat3', ad3, [rnd(g, 1, 16, 4, 4), rnd(g, 1, 16, 4, 4), rnd(g, 1, 32, 4, 4)], 16)
    aa2 = C.Adapt_Add2(); aa2.w.data = torch.tensor([0.8, 1.3])
    block_fixture('adapt_add2', aa2, [rnd(g, 2, 16, 5, 5), rnd(g, 2, 16, 5, 5)], 17)
    block_fixture('adapt_add3', C.Adapt_Add3(16, 16, 32), [rnd(g, 1, 16, 5, 5), rnd(g, 1, 16, 5, 5), rnd(g, 1, 32, 5, 5)], 18)
    # ---- a7 pools ----
    block_fixture('sppf_20', C.SPPF(32, 32, 5), [rnd(g, 1, 32, 20, 20)], 19)
    block_fixture('sppfcspc_12x9', C.SPPFCSPC(32, 32), [rnd(g, 2, 32, 12, 9)], 20)
    block_fixture('spp', C.SPP(32, 32), [rnd(g, 1, 32, 10, 13)], 21)
    block_fixture('sppcspc', C.SPPCSPC(32, 32), [rnd(g, 1, 32, 7, 9)], 22)
    x = rnd(g, 1, 16, 48, 48)
    y1, y2, y3 = [torch.nn.functional.max_pool2d(x, k, 1, k // 2) for k in (5, 9, 13)]
    save('maxpool_cascade_48', in0=x, y1=y1, y2=y2, y3=y3)
    # ---- upsample + concat glue ----
    up = nn.Upsample(None, 2, 'nearest')
    x = rnd(g, 2, 16, 5, 7)
    save('upsample2', in0=x, out=up(x))
    # ---- a8 Detect ----
    det = Y.Detect(nc=4, anchors=[[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]],
                   ch=(16, 32, 64))
    det.stride = torch.tensor([8., 16., 32.])
    det.anchors /= det.stride.view(-1, 1, 1)
    for mi in det.m:
        mi.bias.data.normal_(0, 1, generator=g)
    round_module(det)
    det.eval()
    xs = [rnd(g, 2, 16, 8, 6), rnd(g, 2, 32, 4, 3), rnd(g, 2, 64, 2, 2)]
    with torch.no_grad():
        pred, raw = det([t.clone() for t in xs])
    arrs = {f'in{i}': t for i, t in enumerate(xs)}
    arrs.update({f'raw{i}': t for i, t in enumerate(raw)})
    arrs['out'] = pred
    for k, v in det.state_dict().items():
        arrs['sd/' + k] = v
    save('detect_nc4', **arrs)

    # ---- a9 NMS ----
    nms = R.non_max_suppression
    gn = torch.Generator().manual_seed(7)
    pred = torch.rand(3, 3000, 15, generator=gn)
    pred[..., :2] *= 640
    pred[..., 2:4] = pred[..., 2:4] * 60 + 4
    arrs = {'pred': pred}
    styles = {'detect': dict(conf_thres=0.25, iou_thres=0.45, max_det=1000),
              'val': dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300),
              'agnostic': dict(conf_thres=0.3, iou_thres=0.5, agnostic=True, max_det=50),
              'classes': dict(conf_thres=0.2, iou_thres=0.45, classes=[1, 3, 7], max_det=300)}
    for sname, kw in styles.items():
        for i in range(pred.shape[0]):
            o = nms(pred[i:i + 1].clone(), **kw)[0]
            arrs[f'{sname}_{i}'] = o
    save('nms_random', **arrs)
    # dense overlap set (boxes clustered so that suppression chains are long)
    centers = torch.rand(40, 2, generator=gn) * 500 + 50
    idx = torch.randint(0, 40, (2, 4000), generator=gn)
    pred2 = torch.zeros(2, 4000, 8)
    pred2[..., :2] = centers[idx] + torch.randn(2, 4000, 2, generator=gn) * 6
    pred2[..., 2:4] = 60 + torch.randn(2, 4000, 2, generator=gn).abs() * 10
    pred2[..., 4] = torch.rand(2, 4000, generator=gn)
    pred2[..., 5:] = torch.rand(2, 4000, 3, generator=gn)
    arrs = {'pred': pred2}
    for sname, kw in styles.items():
        if sname == 'classes':
            kw = dict(kw, classes=[0, 2])
        for i in range(2):
            arrs[f'{sname}_{i}'] = nms(pred2[i:i + 1].clone(), **kw)[0]
    save('nms_clustered', **arrs)
    # > max_nms truncation: 7000 rows x 5 classes multi-label = 35k candidates, scores DISTINCT by
    # construction (obj = 1, class scores = a permutation of an arithmetic sequence) because the reference's
    # argsort(descending=True) is unstable, i.e. implementation-defined on ties.
    gn2 = torch.Generator().manual_seed(11)
    pred3 = torch.rand(1, 7000, 10, generator=gn2)
    pred3[..., :2] *= 640
    pred3[..., 2:4] = pred3[..., 2:4] * 40 + 4
    pred3[..., 4] = 1.0
    vals = 0.5 + torch.arange(35000, dtype=torch.float32) * 1e-5
    pred3[..., 5:] = vals[torch.randperm(35000, generator=gn2)].view(1, 7000, 5)
    assert pred3[..., 5:].unique().numel() == 35000
    o = nms(pred3.clone(), conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)[0]
    save('nms_truncate', pred=pred3, val_0=o)

    # ---- model-level (seeded, BN-calibrated, bf16-rounded weights) ----
    import yaml
    for cfg_name, nc, hw in (('yolov5s.yaml', None, 64), ('ablation-ca-scconv-sppfcspc-bifpn.yaml', None, 64)):
        cfg = str(Path('/root/reference/models') / cfg_name)
        m = build_calibrated(R.Model, cfg, seed=0, nc=nc, calib_hw=(128, 128), calib_bs=2)
        x = torch.rand(2, 3, hw, hw, generator=torch.Generator().manual_seed(1)).bfloat16().float()
        with torch.no_grad():
            pred, raw = m(x)
        arrs = {'in0': x, 'out': pred, 'digest': np.array(state_digest(m.state_dict())),
                'strides': m.stride}
        # per-layer statistics of the reference activations (for a layer-by-layer drift report)
        stats = []
        hooks = []
        for layer in m.model:
            hooks.append(layer.register_forward_hook(
                lambda mod, i, o: stats.append([float(o.abs().mean()), float(o.abs().max())]) if torch.is_tensor(o) else None))
        with torch.no_grad():
            m(x)
        for h in hooks:
            h.remove()
        arrs['layer_stats'] = np.array(stats, dtype=np.float32)
        for style, kw in (('detect', dict(conf_thres=0.25, iou_thres=0.45, max_det=1000)),
                          ('val', dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300))):
            for i in range(pred.shape[0]):
                arrs[f'nms_{style}_{i}'] = nms(pred[i:i + 1].clone(), **kw)[0]
        save('model_' + cfg_name.replace('.yaml', ''), **arrs)


if __name__ == '__main__':
    main()
