import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dma_yolo_b200 import ops
from tools.bench_kernels import timeit
dev='cuda'; B=64
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for cin, ho in [(64, 160), (128, 80), (256, 40)]:
    x = ops.empty_nhwc(B, cin, ho, ho, dev).normal_()
    res = ops.empty_nhwc(B, cin, ho, ho, dev).normal_()
    k2 = ops.empty_nhwc(B, cin, ho // 4, ho // 4, dev).normal_()
    pk = ops.pack_conv(torch.randn(cin, cin, 3, 3) / (cin * 9) ** 0.5, stride=1, pad=1, device=dev)
    out = ops.empty_nhwc(B, cin, ho, ho, dev)
    slab = ops.empty_nhwc(B, 2 * cin, ho, ho, dev)
    fl = 2 * B * ho * ho * cin * cin * 9 / 1e9
    cases = {'plain': lambda: ops.conv(x, pk, 1, out=out),
             'linear(no act)': lambda: ops.conv(x, pk, 0, out=out),
             'residual': lambda: ops.conv(x, pk, 1, out=out, residual=res),
             'residual=x': lambda: ops.conv(x, pk, 1, out=out, residual=x),
             'slab out': lambda: ops.conv(x, pk, 1, out=slab[:, :cin]),
             'slab out + residual': lambda: ops.conv(x, pk, 1, out=slab[:, :cin], residual=res),
             'gate': lambda: ops.conv(x, pk, 0, out=out, gate=(x, k2))}
    for name, fn in cases.items():
        ms = timeit(fn, reps=7, flush=flush)
        ms2 = timeit(fn, reps=7, flush=None)
        print(json.dumps(dict(conv=f'{cin}->{cin} k3 @{ho}', case=name, ms_cold=round(ms, 4), ms_warm=round(ms2, 4), tflops_cold=round(fl / ms, 1))), flush=True)
