"""Experiment: run the first layers (stem -> SCConv -> C3) per batch CHUNK so that a chunk's activations stay in the 126 MB L2
between the producer and its consumers.  python tools/exp_chunked.py [last_layer]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dma_yolo_b200 as D  # noqa: E402
from dma_yolo_b200 import ops  # noqa: E402
from dma_yolo_b200.utils.calib import build_calibrated  # noqa: E402

last = int(sys.argv[1]) if len(sys.argv) > 1 else 1
m = build_calibrated('ablation-ca-scconv-sppfcspc-bifpn.yaml', seed=0, calib_hw=(320, 320), calib_bs=4).cuda().eval()
x = torch.rand(64, 3, 640, 640, generator=torch.Generator().manual_seed(1)).cuda()
layers = list(m.model[:last + 1])
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def run(xb):
    t = xb
    for l in layers:
        t = m._run_layer(l, t, True)
    return t


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


with torch.no_grad():
    ref = run(x)
    print('layers 0..%d full batch: %.3f ms' % (last, timed(lambda: run(x))))
    for ch in (32, 16, 8, 4, 2):
        def chunked():
            return [run(x[i:i + ch]) for i in range(0, 64, ch)]
        outs = chunked()
        ok = torch.equal(torch.cat([o for o in outs]), ref)
        print('chunk %2d: %.3f ms  equal=%s' % (ch, timed(chunked), ok))
