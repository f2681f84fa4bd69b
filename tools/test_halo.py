"""Correctness + timing of the halo conv path variants (flags, see include/dmayolo.h)."""
import json, sys
from pathlib import Path
import torch
import torch.nn.functional as F
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dma_yolo_b200 import ops
from tools.bench_kernels import timeit

dev = 'cuda'
def check(n, cin, h, w, cout, flags, gate=False, res=False):
    g = torch.Generator().manual_seed(cin + h)
    x = torch.randn(n, cin, h, w, generator=g).bfloat16().float()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).bfloat16().float()
    pk = ops.pack_conv(wt, stride=1, pad=1, device=dev)
    ref = F.conv2d(x, wt, None, 1, 1)
    ref = ref * torch.sigmoid(ref)
    xa = ops.as_act(x.to(dev))
    try:
        y = ops.conv(xa, pk, 1, flags=flags).float().cpu()
        torch.cuda.synchronize()
    except Exception as e:
        return f'EXC {e}'
    err = (y - ref).abs()
    bad = int((err > 1e-2 + 1e-2 * ref.abs()).sum())
    return f'bad={bad}/{err.numel()} maxerr={float(err.max()):.4f}'

variants = {'im2col': 1, 'halo (resident if fits)': 2, 'halo streamed B': 2 | 4, 'auto': 0}
shapes = [(2, 64, 32, 16, 64), (1, 128, 48, 40, 128), (2, 64, 20, 20, 64), (1, 16, 32, 32, 32), (1, 32, 16, 24, 64), (3, 256, 16, 8, 256)]
for name, fl in variants.items():
    for sh in shapes:
        print(json.dumps(dict(variant=name, shape=sh, result=check(*sh, fl))), flush=True)

flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
B = 64
for cin, cout, ho in [(64, 64, 320), (64, 64, 160), (128, 128, 160), (128, 128, 80), (256, 256, 80), (256, 256, 40), (16, 64, 320)]:
    x = ops.empty_nhwc(B, cin, ho, ho, dev).normal_()
    pk = ops.pack_conv(torch.randn(cout, cin, 3, 3) / (cin * 9) ** 0.5, stride=1, pad=1, device=dev)
    out = ops.empty_nhwc(B, cout, ho, ho, dev)
    for name, fl in variants.items():
        try:
            ms = timeit(lambda: ops.conv(x, pk, 1, out=out, flags=fl), reps=5, flush=flush)
            print(json.dumps(dict(time=name, conv=f'{cin}->{cout} @{ho}', ms=round(ms, 4), tflops=round(2 * B * ho * ho * cout * cin * 9 / ms / 1e9, 1))), flush=True)
        except Exception as e:
            print(json.dumps(dict(time=name, conv=f'{cin}->{cout} @{ho}', err=str(e)[:100])), flush=True)
    del x, out
