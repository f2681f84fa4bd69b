"""Isolated sweep of the n-tile width (dmay_conv_params.block_n) and a few plan flags over the small-M / 1x1 layers of cfg-2:
python tools/sweep_block_n.py   (L2 flushed before every call; see tools/bench_kernels.timeit)"""
import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dma_yolo_b200 import ops
from tools.bench_kernels import timeit
B, dev = 64, 'cuda'
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
#        cin  cout k  ho  mode
convs = [(512, 512, 1, 20, 'plain'), (1024, 1024, 1, 20, 'plain'), (1024, 512, 1, 20, 'plain'), (512, 512, 3, 20, 'plain'),
         (512, 512, 3, 20, 'res'), (256, 256, 3, 20, 'plain'), (1024, 1024, 3, 20, 'plain'), (2048, 1024, 1, 20, 'plain'),
         (4096, 1024, 1, 20, 'plain'), (256, 256, 1, 40, 'plain'), (512, 512, 1, 40, 'plain'), (512, 256, 1, 40, 'plain'),
         (1024, 512, 1, 40, 'plain'), (128, 128, 1, 80, 'plain'), (256, 256, 1, 80, 'plain'), (64, 64, 1, 160, 'plain'),
         (128, 128, 1, 160, 'plain'), (256, 256, 3, 40, 'res'), (256, 256, 3, 40, 'plain')]
for cin, cout, k, ho, mode in convs:
    x = ops.empty_nhwc(B, cin, ho, ho, dev).normal_()
    pk = ops.pack_conv(torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5, stride=1, pad=k // 2, device=dev)
    out = ops.empty_nhwc(B, cout, ho, ho, dev)
    kw = {}
    if mode == 'res':
        kw['residual'] = ops.empty_nhwc(B, cout, ho, ho, dev).normal_()
    row = dict(conv=f'{cin}->{cout} k{k} @{ho} {mode}')
    ref = None
    for name, bn, fl in (('default', 0, 0), ('bn64', 64, 0), ('bn128', 128, 0), ('bn256', 256, 0), ('bn128_nopair', 128, 512),
                         ('bn256_nopair', 256, 512), ('bn128_pair', 128, 256), ('epi16', 0, 16), ('epi8', 0, 8)):
        if bn > cout:
            continue
        try:
            ms = timeit(lambda: ops.conv(x, pk, 1, out=out, block_n=bn, flags=fl, **kw), reps=5, flush=flush)
        except Exception as e:  # unsupported combination
            row[name] = 'n/a'
            continue
        row[name] = round(ms * 1000, 1)
        if ref is None:
            ref = out.clone()
        elif not torch.equal(ref, out):
            row[name + '_differs'] = True
    print(json.dumps(row), flush=True)
