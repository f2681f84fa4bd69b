"""A/B of conv tuning flags per epilogue mode:  python tools/bench_conv_modes.py "0,4,20" [shape ...]
   shapes as cin:cout:ho (3x3 s1); modes plain / residual / gate; one line per (shape, mode, flags)."""
import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dma_yolo_b200 import ops
from tools.bench_kernels import timeit
flags_list = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else '0').split(',')]
shapes = [tuple(int(v) for v in a.split(':')) for a in sys.argv[2:]] or [(128, 128, 80), (128, 128, 160), (64, 64, 320), (64, 64, 160)]
B, dev = 64, 'cuda'
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for cin, cout, ho in shapes:
    x = ops.empty_nhwc(B, cin, ho, ho, dev).normal_()
    res = ops.empty_nhwc(B, cout, ho, ho, dev).normal_()
    k2 = ops.empty_nhwc(B, cout, ho // 4, ho // 4, dev).normal_()
    pk = ops.pack_conv(torch.randn(cout, cin, 3, 3) / (cin * 9) ** 0.5, stride=1, pad=1, device=dev)
    out = ops.empty_nhwc(B, cout, ho, ho, dev)
    for mode in ('plain', 'residual', 'gate'):
        if mode == 'gate' and cin != cout:
            continue
        for fl in flags_list:
            kw = dict(residual=res) if mode == 'residual' else dict(gate=(x, k2)) if mode == 'gate' else {}
            act = 0 if mode == 'gate' else 1
            try:
                ms = timeit(lambda: ops.conv(x, pk, act, out=out, flags=fl, **kw), reps=7, flush=flush)
            except Exception as e:   # a flag combination may not fit shared memory
                print(json.dumps(dict(conv=f'{cin}->{cout} @{ho}', mode=mode, flags=fl, error=str(e)[:60])), flush=True)
                continue
            print(json.dumps(dict(conv=f'{cin}->{cout} @{ho}', mode=mode, flags=fl, ms=round(ms, 4),
                                  tflops=round(2 * B * ho * ho * cout * cin * 9 / ms / 1e9, 1))), flush=True)
    del x, res, k2, out
