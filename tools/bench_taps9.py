"""A/B of the one-block nine-tap MMA issue (flags bit16 = old per-tap loop) on the resident-weight 3x3 layers, plain /
residual / gate epilogues: python tools/bench_taps9.py"""
import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dma_yolo_b200 import ops
from tools.bench_kernels import timeit
B, dev = 64, 'cuda'
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
convs = [(128, 128, 160, 'plain'), (128, 128, 160, 'gate'), (128, 128, 80, 'plain'), (128, 128, 80, 'res'), (256, 256, 40, 'res'), (256, 256, 80, 'gate'), (512, 512, 20, 'res'), (64, 64, 160, 'res'), (64, 64, 320, 'gate'), (16, 64, 320, 'plain'), (64, 64, 320, 'plain'), (64, 64, 320, 'gate'), (64, 64, 160, 'res'), (64, 64, 160, 'plain'), (64, 64, 80, 'plain'),
         (32, 32, 320, 'plain'), (128, 128, 160, 'plain')]
for cin, cout, ho, mode in convs:
    x = ops.empty_nhwc(B, cin, ho, ho, dev).normal_()
    pk = ops.pack_conv(torch.randn(cout, cin, 3, 3) / (cin * 9) ** 0.5, stride=1, pad=1, device=dev)
    out = ops.empty_nhwc(B, cout, ho, ho, dev)
    kw = {}
    if mode == 'res':
        kw['residual'] = ops.empty_nhwc(B, cout, ho, ho, dev).normal_()
    if mode == 'gate':
        kw['gate'] = (x, ops.empty_nhwc(B, cout, ho // 4, ho // 4, dev).normal_())
    act = 0 if mode == 'gate' else 1
    row = dict(conv=f'{cin}->{cout} k3 @{ho} {mode}')
    outs = {}
    for name, fl in (('loop', int(sys.argv[1]) if len(sys.argv) > 1 else 65536), ('taps9', 0)):
        ms = timeit(lambda: ops.conv(x, pk, act, out=out, flags=fl, **kw), reps=7, flush=flush)
        row[name + '_ms'] = round(ms, 4)
        row[name + '_tflops'] = round(2 * B * ho * ho * cout * cin * 9 / ms / 1e9, 1)
        outs[name] = out.clone()
    row['identical'] = bool(torch.equal(outs['loop'], outs['taps9']))
    print(json.dumps(row), flush=True)
