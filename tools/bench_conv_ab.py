"""A/B conv timing: python tools/bench_conv_ab.py [block_n_flag]   (DMAY_SO selects the library)."""
import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dma_yolo_b200 import ops
from tools.bench_kernels import timeit
flag = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B, dev = 64, 'cuda'
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
convs = [(64, 64, 3, 1, 320), (64, 64, 3, 1, 160), (64, 64, 1, 1, 160), (128, 128, 3, 1, 160), (128, 128, 3, 1, 80), (128, 128, 1, 1, 80),
         (256, 256, 3, 1, 80), (256, 256, 3, 1, 40), (256, 256, 1, 1, 40), (256, 256, 1, 1, 80), (512, 512, 1, 1, 40), (512, 512, 1, 1, 20), (1024, 256, 1, 1, 40), (128, 256, 1, 1, 80), (512, 512, 3, 1, 40), (512, 512, 3, 1, 20),
         (1024, 1024, 3, 1, 20), (4096, 1024, 1, 1, 20), (1024, 1024, 1, 1, 20)]
for cin, cout, k, s, ho in convs:
    x = ops.empty_nhwc(B, cin, ho * s, ho * s, dev).normal_()
    pk = ops.pack_conv(torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5, stride=s, pad=k // 2, device=dev)
    out = ops.empty_nhwc(B, cout, ho, ho, dev)
    ms = timeit(lambda: ops.conv(x, pk, 1, out=out, block_n=flag), reps=7, flush=flush)
    print(json.dumps(dict(conv=f'{cin}->{cout} k{k} @{ho}', ms=round(ms, 4), tflops=round(2 * B * ho * ho * cout * cin * k * k / ms / 1e9, 1))), flush=True)
    del x, out
