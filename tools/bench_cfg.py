"""Throughput of any model config on the kernel path (forward + fused decode/filter + NMS), device-timed.

    python tools/bench_cfg.py --cfg yolov5l-ca-sppfcspc-bifpn-scconv.yaml --imgsz 1536 --bs 8 [--style val|detect]

Same methodology as bench.py (CUDA events, >= 3 warm-up steps, inputs resident in HBM) for the BASELINE.json configs
that are parity cases rather than the headline (cfg-3, cfg-4b); prints one JSON line."""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cfg', required=True)
    ap.add_argument('--imgsz', type=int, default=640)
    ap.add_argument('--bs', type=int, default=8)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--style', default='val')
    a = ap.parse_args()
    import dma_yolo_b200 as D
    from dma_yolo_b200.utils.calib import build_calibrated
    m = build_calibrated(a.cfg, seed=0, calib_hw=(320, 320), calib_bs=4).cuda().eval()   # SURVEY.md F5, as bench.py
    x = torch.rand(a.bs, 3, a.imgsz, a.imgsz, generator=torch.Generator().manual_seed(1)).cuda()
    kw = dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300) if a.style == 'val' else \
        dict(conf_thres=0.25, iou_thres=0.45, max_det=1000)
    n0 = D.launch_count()

    def step():
        with torch.no_grad():
            pred, _ = m(x)
            return D.non_max_suppression(pred, **kw)
    for _ in range(a.warmup):
        dets = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        dets = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps(dict(cfg=a.cfg, imgsz=a.imgsz, batch=a.bs, style=a.style, ms_per_step=round(ms, 3),
                          img_per_s=round(a.bs / ms * 1e3, 1), detections=[int(len(d)) for d in dets][:8],
                          launches_per_step=(D.launch_count() - n0) // (a.steps + a.warmup))), flush=True)


if __name__ == '__main__':
    main()
