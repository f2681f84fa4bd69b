"""How finely can mAP@0.5:0.95 be resolved on the conditioned checkpoints?  (test infrastructure; CPU, ~2 min)

    python -m oracle.map_sensitivity

For both conditioned checkpoints (tests/golden/conditioned_*.pt): the fp32 oracle chain (== the executed reference), the
bf16-storage oracle, and the bf16-storage oracle with a 1e-5 RELATIVE perturbation of the decoded prediction.  Measured:
    ablation  fp32 0.953702   bf16 +9.0e-5   bf16 + 1e-5 noise +9.0e-5
    c3caspd   fp32 0.912103   bf16 -5.5e-4   bf16 + 1e-5 noise +9.3e-4      (a 1e-5 perturbation moves the metric by 1.5e-3)
One detection that NMS keeps or drops changes TP/FP at up to ten IoU levels: up to 10 / (n_labels(class) * 10 * nc) = 7.8e-3.
"""
import sys, json, torch, numpy as np
sys.path.insert(0, '.')
from pathlib import Path
from oracle import blocks as O, nms as ON, metrics as OM, synth
torch.set_num_threads(16)
gold = Path('tests/golden')
for tag in ('ablation', 'c3caspd'):
    ck = torch.load(gold / f'conditioned_{tag}.pt', map_location='cpu')
    info = json.load(open(gold / f'conditioned_{tag}.json'))
    sd = {k: (v.float() if v.is_floating_point() else v) for k, v in ck['state_dict'].items()}
    S = info['size']
    kw = dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)
    for mode in ('fp32', 'bf16', 'bf16+noise'):
        dets, labels = [], []
        for b in range(info['val_batches']):
            im, tg = synth.make_batch(info['val_seed'] + b, info['val_b'], S)
            x = torch.from_numpy(im).float() / 255
            with torch.no_grad():
                if mode == 'fp32':
                    pred, _, _ = O.forward_model(ck['cfg'], sd, x, ck['stride'])
                else:
                    with O.bf16_storage():
                        pred, _, _ = O.forward_model(ck['cfg'], sd, x.bfloat16().float(), ck['stride'])
                    if mode == 'bf16+noise':   # a 1e-5 relative perturbation of the decoded prediction
                        g = torch.Generator().manual_seed(b)
                        pred = pred * (1 + 1e-5 * torch.randn(pred.shape, generator=g))
            for i in range(im.shape[0]):
                dets.append(ON.non_max_suppression(pred[i:i + 1].numpy(), **kw)[0])
                labels.append(tg[tg[:, 0] == i][:, 1:])
        mp, mr, map50, map_ = OM.evaluate(dets, labels, (S, S))
        print(tag, mode, 'P %.5f R %.5f mAP50 %.5f mAP %.6f  delta vs reference %.2e' % (mp, mr, map50, map_, map_ - info['map']), flush=True)
