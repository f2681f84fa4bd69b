"""Produce a CONDITIONED (trained) checkpoint + its golden val statistics by EXECUTING the unmodified reference in the
build container: reference `Model` + reference `ComputeLoss` (utils/loss.py:135) trained for a few hundred CPU steps on
the seeded synthetic set of oracle/synth.py, weights rounded to bf16, then the reference's own `val.run` (val.py:87)
on a seeded synthetic val set -> tests/golden/conditioned_<tag>.{pt,json}.

    python -m oracle.train_conditioned --cfg ablation-ca-scconv-sppfcspc-bifpn --tag ablation --steps 900

Why: an untrained BN-calibrated net amplifies one-ulp differences ~1.3x per layer, so whole-path
criteria (free-running activations, mAP within 1e-4) cannot be judged on it.  The width/depth multiples are reduced
(0.25 / 0.33) so that the bf16 state_dict stays a small committed fixture; every custom block of the config is kept.
Test infrastructure only: nothing in the product imports this.
"""
from __future__ import annotations

import argparse
import json
import math
import sys
import time
from pathlib import Path

import numpy as np
import torch
import yaml

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import synth  # noqa: E402

GOLD = ROOT / 'tests' / 'golden'
VAL_SEED, VAL_B, VAL_BATCHES = 9000, 8, 4


def reduced_cfg(name: str, width=0.25, depth=0.33, nc=synth.NC):
    d = yaml.safe_load(open(ROOT / 'dma_yolo_b200' / 'models' / f'{name}.yaml'))
    d['width_multiple'], d['depth_multiple'], d['nc'] = width, depth, nc
    return d


def val_loader(S):
    out = []
    for b in range(VAL_BATCHES):
        im, tg = synth.make_batch(VAL_SEED + b, VAL_B, S)
        shapes = [((S, S), ((1.0, 1.0), (0.0, 0.0)))] * VAL_B
        out.append((torch.from_numpy(im), torch.from_numpy(tg), [f'synth_{b}_{i}.jpg' for i in range(VAL_B)], shapes))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cfg', default='ablation-ca-scconv-sppfcspc-bifpn')
    ap.add_argument('--tag', default='ablation')
    ap.add_argument('--steps', type=int, default=900)
    ap.add_argument('--bs', type=int, default=8)
    ap.add_argument('--size', type=int, default=640)
    ap.add_argument('--train-size', type=int, default=0)
    ap.add_argument('--lr', type=float, default=2e-3)
    ap.add_argument('--width', type=float, default=0.25)
    ap.add_argument('--depth', type=float, default=0.33)
    ap.add_argument('--anchors', default='', help="'p2': replace a placeholder anchor COUNT by four real anchor triples (P2..P5)")
    args = ap.parse_args()
    from oracle import refshim
    R = refshim.load()
    from utils.loss import ComputeLoss
    import val as V

    S = args.size
    St = args.train_size or S
    cfgd = reduced_cfg(args.cfg, args.width, args.depth)
    if args.anchors == 'p2':   # C3CASPD.yaml / spdconv.yaml say `anchors: 4` (placeholders 0..7: zero-size boxes, nothing to train on)
        cfgd['anchors'] = [[5, 6, 8, 14, 15, 11], [10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]
    torch.manual_seed(0)
    torch.set_num_threads(8)
    model = R.Model(cfgd, ch=3, nc=synth.NC)
    hyp = yaml.safe_load(open('/root/reference/data/hyps/hyp.scratch.yaml'))
    nl = model.model[-1].nl
    hyp['box'] *= 3 / nl
    hyp['cls'] *= synth.NC / 80 * 3 / nl
    hyp['obj'] *= (St / 640) ** 2 * 3 / nl
    hyp['label_smoothing'] = 0.0
    model.nc, model.hyp, model.gr = synth.NC, hyp, 1.0
    model.names = ['red_rect', 'green_rect', 'blue_ellipse', 'yellow_ellipse']
    loss_fn = ComputeLoss(model)
    opt = torch.optim.Adam(model.parameters(), lr=args.lr, betas=(0.9, 0.999))
    model.train()
    t0 = time.time()
    for step in range(args.steps):
        lr = args.lr * (min(1.0, (step + 1) / 50) * 0.5 * (1 + math.cos(math.pi * step / args.steps)) + 0.01)
        for gparam in opt.param_groups:
            gparam['lr'] = lr
        im, tg = synth.make_batch(100000 + step, args.bs, St)
        x = torch.from_numpy(im).float() / 255
        pred = model(x)
        loss, items = loss_fn(pred, torch.from_numpy(tg))
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
        opt.step()
        if step % 25 == 0 or step == args.steps - 1:
            print(f'step {step} lr {lr:.5f} loss {float(loss):.4f} box/obj/cls {items.tolist()} {time.time() - t0:.0f}s', flush=True)
    model.eval()
    # identical weights on both sides of every parity test: every float parameter / buffer is a bf16 value
    for p in model.parameters():
        p.data = p.data.bfloat16().float()
    for _, b in model.named_buffers():
        if b.is_floating_point():
            b.data = b.data.bfloat16().float()
    sd = {k: (v.bfloat16() if v.is_floating_point() else v.clone()) for k, v in model.state_dict().items()}
    loader = val_loader(S)
    res, maps, _ = V.run({'nc': synth.NC, 'val': 'synthetic', 'names': model.names}, model=model,
                         dataloader=[(a, b.clone(), c, d) for a, b, c, d in loader], batch_size=VAL_B, imgsz=S, plots=False)
    mp, mr, map50, map_ = [float(v) for v in res[:4]]
    print(f'reference val.run: P {mp:.4f} R {mr:.4f} mAP50 {map50:.4f} mAP {map_:.6f}')
    # the reference's detections of the first val batch (diagnostics for the parity test)
    with torch.no_grad():
        pred = model(loader[0][0].float() / 255)[0]
    dets = []
    for i in range(VAL_B):
        dets.append(R.non_max_suppression(pred[i:i + 1], 0.001, 0.6, multi_label=True)[0].numpy())
    torch.save({'state_dict': sd, 'cfg': cfgd, 'stride': model.stride.tolist()}, GOLD / f'conditioned_{args.tag}.pt')
    json.dump(dict(cfg=args.cfg, width=args.width, depth=args.depth, size=S, steps=args.steps, val_seed=VAL_SEED, val_b=VAL_B,
                   val_batches=VAL_BATCHES, mp=mp, mr=mr, map50=map50, map=map_, maps=[float(v) for v in maps],
                   n_dets_batch0=[int(len(d)) for d in dets],
                   top_conf_batch0=[[float(v) for v in d[:8, 4]] for d in dets]),
              open(GOLD / f'conditioned_{args.tag}.json', 'w'), indent=1)
    np.savez_compressed(GOLD / f'conditioned_{args.tag}_dets.npz', **{f'det_{i}': d for i, d in enumerate(dets)})
    print('wrote', GOLD / f'conditioned_{args.tag}.pt', (GOLD / f'conditioned_{args.tag}.pt').stat().st_size >> 10, 'KiB')


if __name__ == '__main__':
    main()
