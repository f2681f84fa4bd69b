"""Run ONE kernel configuration a few times (for `ncu`): python tools/prof_one.py conv 128 128 3 1 80 [bs]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dma_yolo_b200 import ops  # noqa: E402

kind = sys.argv[1]
dev = 'cuda'
if kind == 'conv':
    cin, cout, k, s, ho = map(int, sys.argv[2:7])
    B = int(sys.argv[7]) if len(sys.argv) > 7 else 64
    x = ops.empty_nhwc(B, cin, ho * s, ho * s, dev).normal_()
    pk = ops.pack_conv(torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5, stride=s, pad=k // 2, device=dev)
    out = ops.empty_nhwc(B, cout, ho, ho, dev)
    for _ in range(3):
        ops.conv(x, pk, 1, out=out)
elif kind == 'convres':
    # Bottleneck cv2 with the residual add in the epilogue: python tools/prof_one.py convres 128 80
    c, s_ = map(int, sys.argv[2:4])
    B = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    x = ops.empty_nhwc(B, c, s_, s_, dev).normal_()
    r = ops.empty_nhwc(B, c, s_, s_, dev).normal_()
    pk = ops.pack_conv(torch.randn(c, c, 3, 3) / (c * 9) ** 0.5, stride=1, pad=1, device=dev)
    out = ops.empty_nhwc(B, c, s_, s_, dev)
    for _ in range(3):
        ops.conv(x, pk, 1, out=out, residual=r)
elif kind == 'convgate':
    # SCConv k3 with the gate in the epilogue: k3(x) * sigmoid(x + up(k2))   python tools/prof_one.py convgate 64 320
    c, s = map(int, sys.argv[2:4])
    B = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    x = ops.empty_nhwc(B, c, s, s, dev).normal_()
    k2 = ops.empty_nhwc(B, c, s // 4, s // 4, dev).normal_()
    pk = ops.pack_conv(torch.randn(c, c, 3, 3) / (c * 9) ** 0.5, stride=1, pad=1, device=dev)
    out = ops.empty_nhwc(B, c, s, s, dev)
    for _ in range(3):
        ops.conv(x, pk, 0, out=out, gate=(x, k2))
elif kind == 'gate':
    c, s = map(int, sys.argv[2:4])
    B = 64
    x, k3, k2, out = (ops.empty_nhwc(B, c, s, s, dev).normal_(), ops.empty_nhwc(B, c, s, s, dev).normal_(),
                      ops.empty_nhwc(B, c, s // 4, s // 4, dev).normal_(), ops.empty_nhwc(B, c, s, s, dev))
    for _ in range(3):
        ops.scconv_gate(x, k3, k2, out=out)
elif kind == 'nms5':
    # BASELINE cfg-5: decode-less non_max_suppression on rand(256, 25200, 15), val style; second call profiled
    import dma_yolo_b200 as D
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    pred = torch.rand(B, 25200, 15, generator=torch.Generator().manual_seed(1))
    pred[..., :2] *= 640
    pred[..., 2:4] = pred[..., 2:4] * 60 + 4
    pred = pred.to(dev)
    for it in range(3):
        if it == 2:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        D.non_max_suppression(pred, 0.001, 0.6, multi_label=True, max_det=300)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
elif kind == 'model':
    # whole hot path (forward + fused decode/filter + NMS) once warm, once profiled-range: for `ncu -k regex:...`
    import dma_yolo_b200 as D
    from dma_yolo_b200.utils.calib import build_calibrated
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    cfg = sys.argv[3] if len(sys.argv) > 3 else 'ablation-ca-scconv-sppfcspc-bifpn.yaml'
    S = int(sys.argv[4]) if len(sys.argv) > 4 else 640
    m = build_calibrated(cfg, seed=0, calib_hw=(320, 320), calib_bs=4).to(dev).eval()   # SURVEY.md F5, as bench.py
    x = torch.rand(B, 3, S, S, generator=torch.Generator().manual_seed(1)).to(dev)
    with torch.no_grad():
        for it in range(2):
            if it == 1:
                torch.cuda.synchronize()
                torch.cuda.profiler.start()
            pred, _ = m(x)
            D.non_max_suppression(pred, 0.001, 0.6, multi_label=True, max_det=300)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
torch.cuda.synchronize()
print('ok')
