"""Where does the fused filter's time go inside the network?  Event-timed dmay_nms_filter_fused (rows kernel + scan + gather)
(a) right behind the forward that wrote the logits, (b) alone on the same logits after a sync and an L2 flush.
python tools/exp_filter_timing.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dma_yolo_b200 import ops  # noqa: E402
from dma_yolo_b200.utils.calib import build_calibrated  # noqa: E402

m = build_calibrated('ablation-ca-scconv-sppfcspc-bifpn.yaml', seed=0, calib_hw=(320, 320), calib_bs=4).cuda().eval()
x = torch.rand(64, 3, 640, 640, generator=torch.Generator().manual_seed(1)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def filt(pred):
    return ops._fused_candidates(pred._levels, pred._na, pred._no - 5, 0.001, True, None, max_nms=30000)


def ev():
    return torch.cuda.Event(enable_timing=True)


with torch.no_grad():
    for _ in range(3):
        pred, _ = m(x)
        filt(pred)
    a_ms, b_ms = [], []
    for _ in range(10):
        pred, _ = m(x)
        e0, e1 = ev(), ev()
        e0.record()
        filt(pred)           # includes its allocations / zeroing and the sizing sync at the end
        e1.record()
        torch.cuda.synchronize()
        a_ms.append(e0.elapsed_time(e1))
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        filt(pred)
        e1.record()
        torch.cuda.synchronize()
        b_ms.append(e0.elapsed_time(e1))
    print('behind the forward: median %.3f ms   alone after sync + flush: median %.3f ms' % (sorted(a_ms)[5], sorted(b_ms)[5]))
