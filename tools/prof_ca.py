"""One CoordAtt call at the cfg-2 shape (64 x 1024 x 20 x 20), for ncu:  python tools/prof_ca.py [N C H W]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from dma_yolo_b200 import ops, _lib
from dma_yolo_b200.models import common as C

n, c, h, w = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (64, 1024, 20, 20)
torch.manual_seed(0)
m = C.CoorAttention(c, c).eval()
x = ops.as_act(torch.randn(n, c, h, w).cuda())
pk = ops.pack_coordatt(m.conv1, m.bn1, m.conv_h, m.conv_w, 'cuda')
out = ops.empty_nhwc(n, c, h, w, 'cuda')
l0 = _lib.launch_count()
for _ in range(3):
    ops.coordatt(x, pk, out=out)
torch.cuda.synchronize()
print('launches per call', (_lib.launch_count() - l0) / 3)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    ops.coordatt(x, pk, out=out)
    with torch.cuda.graph(g, stream=s):
        for _ in range(20):
            ops.coordatt(x, pk, out=out)
    g.replay()
    ev[0].record(s)
    g.replay()
    ev[1].record(s)
torch.cuda.synchronize()
us = ev[0].elapsed_time(ev[1]) * 1000 / 20
print('us per call (back-to-back, L2-warm)', round(us, 2), 'GB/s', round(2 * n * c * h * w * 2 / us / 1e3, 1))
