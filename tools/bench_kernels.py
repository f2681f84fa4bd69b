"""Per-kernel timing at cfg-2 shapes (bs 64, 640x640) with CUDA events; prints achieved GB/s or TFLOP/s.

    python tools/bench_kernels.py [--bs 64] [--quick]

Algorithmic bytes/flops follow SURVEY.md 8a/8d: unique input + output bytes (bf16 = 2 B), 2*M*N*K flops.
Every timed region is preceded by an L2 flush (writing a 256 MiB scratch buffer) unless --no-flush.
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from dma_yolo_b200 import ops  # noqa: E402

PEAKS = {'hbm_gbs': 6555.2, 'bf16_tflops': 1664.3}
try:
    PEAKS.update(json.load(open(ROOT / 'MEASURED_PEAKS.json')))
except Exception:
    pass


def _graph_ms(body, reps):
    """CUDA-graph `reps` repetitions of body() and time one replay (no CPU launch overhead inside the region)."""
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        body()                                   # warm-up outside capture (lazy attribute setting, allocator)
    torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            body()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[1]


def timeit(fn, reps=10, flush=None):
    """Device time of fn() alone.  With `flush`: every call is preceded by a write of a buffer larger than L2, the
    [flush, fn] x reps sequence is replayed as ONE CUDA graph (so no Python / launch overhead sits in the timed
    region — with eager launches a 30 us kernel was measured at 100+ us) and the flush-only graph is subtracted."""
    fn()
    torch.cuda.synchronize()
    reps = max(reps, 10)
    try:
        if flush is not None:
            both = _graph_ms(lambda: (flush.fill_(1.0), fn()), reps)
            only = _graph_ms(lambda: flush.fill_(1.0), reps)
            return max(both - only, 1e-6) / reps
        return _graph_ms(fn, reps) / reps
    except Exception as e:  # capture not possible for this op: eager fallback
        print(json.dumps(dict(note=f'graph capture failed ({type(e).__name__}), eager timing')), flush=True)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            if flush is not None:
                flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--bs', type=int, default=64)
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--no-flush', action='store_true')
    ap.add_argument('--out', default=str(ROOT / 'gpurun_out' / 'bench_kernels.json'))
    a = ap.parse_args()
    dev = 'cuda'
    B = a.bs
    flush = None if a.no_flush else torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    rows = []

    def rec(name, ms, gbytes=None, gflop=None):
        r = dict(kernel=name, ms=round(ms, 4))
        if gbytes is not None:
            r['GB/s'] = round(gbytes / ms * 1e3, 1)
            r['frac_hbm'] = round(r['GB/s'] / PEAKS['hbm_gbs'], 3)
        if gflop is not None:
            r['TFLOP/s'] = round(gflop / ms, 1)
            r['frac_tensor'] = round(r['TFLOP/s'] / PEAKS['bf16_tflops'], 3)
        rows.append(r)
        print(json.dumps(r), flush=True)

    act = lambda c, h, w: ops.empty_nhwc(B, c, h, w, dev).normal_()

    # copy calibration
    src = torch.empty(1 << 30, device=dev, dtype=torch.uint8)
    dst = torch.empty_like(src)
    rec('copy_1GiB(ours)', timeit(lambda: ops.device_copy(dst, src), flush=flush), gbytes=2 * src.numel() / 1e9)
    rec('copy_1GiB(torch)', timeit(lambda: dst.copy_(src), flush=flush), gbytes=2 * src.numel() / 1e9)
    del src, dst

    # a5 SCConv gate, 4 layers
    for c, s in ((64, 320), (128, 160), (256, 80), (512, 40)):
        x, k3, k2 = act(c, s, s), act(c, s, s), act(c, s // 4, s // 4)
        out = act(c, s, s)
        el = B * c * s * s
        rec(f'scconv_gate c{c}@{s}', timeit(lambda: ops.scconv_gate(x, k3, k2, out=out), flush=flush), gbytes=(3 + 1 / 16) * el * 2 / 1e9)
        po = act(c, s // 4, s // 4)
        rec(f'avgpool4 c{c}@{s}', timeit(lambda: ops.avgpool(x, 4, out=po), flush=flush), gbytes=(1 + 1 / 16) * el * 2 / 1e9)
        del x, k3, k2, out, po
    # a6 AdConcat (+fused upsample)
    for (ca, cb, s, up) in ((512, 512, 40, True), (256, 256, 80, True), (512, 1024, 20, False)):
        xa = act(ca, s // 2 if up else s, s // 2 if up else s)
        xb = act(cb, s, s)
        out = act(ca + cb, s, s)
        el_out = B * (ca + cb) * s * s
        el_in = B * (ca * (s * s // 4 if up else s * s) + cb * s * s)
        rec(f'adconcat2 {ca}+{cb}@{s} up={up}', timeit(lambda: ops.adconcat([ops.Up(xa, 1) if up else xa, xb], (0.5, 0.5), out=out), flush=flush),
            gbytes=(el_in + el_out) * 2 / 1e9)
        del xa, xb, out
    x3 = [act(256, 40, 40), act(256, 40, 40), act(512, 40, 40)]
    out = act(1024, 40, 40)
    rec('adconcat3 256+256+512@40', timeit(lambda: ops.adconcat(x3, (0.33, 0.33, 0.33), out=out), flush=flush), gbytes=2 * B * 1024 * 1600 * 2 / 1e9)
    del x3, out
    # a7 SPPF pool cascade
    slab = act(4096, 20, 20)
    rec('sppf_pool3 c1024@20', timeit(lambda: ops.sppf_pool3(slab[:, :1024], slab[:, 1024:2048], slab[:, 2048:3072], slab[:, 3072:], 5), flush=flush),
        gbytes=4 * B * 1024 * 400 * 2 / 1e9)
    del slab
    # a3 CoordAtt
    import torch.nn as nn
    c = 1024
    conv1, bn1, ch, cw = nn.Conv2d(c, 32, 1), nn.BatchNorm2d(32).eval(), nn.Conv2d(32, c, 1), nn.Conv2d(32, c, 1)
    pk = ops.pack_coordatt(conv1, bn1, ch, cw, dev)
    x, out = act(c, 20, 20), act(c, 20, 20)
    rec('coordatt c1024@20', timeit(lambda: ops.coordatt(x, pk, out=out), flush=flush), gbytes=2 * B * c * 400 * 2 / 1e9)
    del x, out
    # a3 at the cfg-4b sizes (C3CASPD.yaml, 1280x1280: CoordAtt inside the CABottlenecks of its four C3CA stages), batch 32
    for c, hw in ((64, 320), (128, 160), (256, 80), (512, 40)):
        cm = max(8, c // 32)
        conv1, bn1, ch, cw = nn.Conv2d(c, cm, 1), nn.BatchNorm2d(cm).eval(), nn.Conv2d(cm, c, 1), nn.Conv2d(cm, c, 1)
        pk = ops.pack_coordatt(conv1, bn1, ch, cw, dev)
        b4 = max(1, B // 2)
        x, out = ops.empty_nhwc(b4, c, hw, hw, dev).normal_(), ops.empty_nhwc(b4, c, hw, hw, dev)
        rec(f'coordatt c{c}@{hw} (cfg-4b, batch {b4})', timeit(lambda: ops.coordatt(x, pk, out=out), flush=flush),
            gbytes=2 * b4 * c * hw * hw * 2 / 1e9)
        del x, out
    # a4 SPD (cfg-4 shape scaled to this batch)
    x = act(64, 320, 320)
    out = act(256, 160, 160)
    rec('spd c64@320', timeit(lambda: ops.spd(x, out=out), flush=flush), gbytes=2 * B * 64 * 320 * 320 * 2 / 1e9)
    del x, out
    # input prep
    img = torch.rand(B, 3, 640, 640, device=dev)
    rec('input_prep f32->spd bf16', timeit(lambda: ops.input_prep(img, True, 16), flush=flush), gbytes=(B * 3 * 640 * 640 * 4 + B * 320 * 320 * 16 * 2) / 1e9)
    del img

    # a1 convs: the 43 shapes would take long; a representative set (Appendix A) incl. the heaviest
    convs = [  # cin, cout, k, s, hw(out)
        (64, 64, 3, 1, 320), (64, 128, 3, 2, 160), (64, 64, 1, 1, 160), (128, 128, 3, 1, 160), (128, 128, 1, 1, 80),
        (128, 128, 3, 1, 80), (256, 256, 3, 1, 80), (256, 256, 1, 1, 40), (256, 256, 3, 1, 40), (512, 512, 3, 1, 40),
        (512, 512, 1, 1, 20), (512, 512, 3, 1, 20), (1024, 1024, 3, 1, 20), (4096, 1024, 1, 1, 20), (1024, 1024, 1, 1, 20),
        (512, 1024, 3, 2, 20), (256, 255, 1, 1, 80),
    ]
    if a.quick:
        convs = convs[:6]
    for cin, cout, k, s, ho in convs:
        hi = ho * s
        x = act(cin, hi, hi)
        w = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5
        pk = ops.pack_conv(w, stride=s, pad=k // 2, device=dev)
        f32 = cout == 255
        out = ops.empty_nhwc(B, cout, ho, ho, dev, torch.float32 if f32 else torch.bfloat16, c_alloc=ops.round_up(cout, 8))
        gflop = 2 * B * ho * ho * cout * cin * k * k / 1e9
        gb = (B * hi * hi * cin * 2 + B * ho * ho * cout * (4 if f32 else 2) + cout * cin * k * k * 2) / 1e9
        ms = timeit(lambda: ops.conv(x, pk, 0 if f32 else 1, out=None if f32 else out, out_fp32=f32), reps=5, flush=flush)
        rec(f'conv {cin}->{cout} k{k}s{s} @{ho}', ms, gbytes=gb, gflop=gflop)
        del x, out
    # stem
    img = torch.rand(B, 3, 640, 640, device=dev)
    pk = ops.pack_conv(torch.randn(64, 3, 6, 6) / 10, stride=2, pad=2, device=dev)
    xs = ops.input_prep(img, True, 16)
    out = act(64, 320, 320)
    pk2 = ops.ConvPack(**{**pk.__dict__, 'stem_spd': False, 'cin': 16})
    rec('conv stem (spd 16ch k3) ->64 @320', timeit(lambda: ops.conv(xs, pk2, 1, out=out), reps=5, flush=flush),
        gbytes=(xs.numel() * 2 + out.numel() * 2) / 1e9, gflop=2 * B * 320 * 320 * 64 * 108 / 1e9)
    Path(a.out).parent.mkdir(exist_ok=True)
    json.dump(rows, open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
