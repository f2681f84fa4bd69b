"""Headline benchmark: images/s of the detection forward hot path (forward + Detect decode + NMS).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config cfg2|cfg3|cfg4a|cfg4b|cfg5]

Default workload (BASELINE.json configs[1], "cfg2"): ablation-ca-scconv-sppfcspc-bifpn.yaml, 640x640, batch 64 per
GPU, bf16 activations, synthetic images, seeded BN-calibrated random-init weights, val-style NMS
(conf 0.001, iou 0.6, multi_label, max_det 300 — val.py:235).  --config selects the other BASELINE.json configs
(parity cases, measured the same way: cfg3 = yolov5l-ca-sppfcspc-bifpn-scconv 1536x1536 batch 8/GPU, cfg4a = spdconv
1280x1280 batch 32, cfg4b = C3CASPD 1280x1280 batch 32, cfg5 = NMS stress 256 x 25,200 x 15, decode-less).

One "step" = one batch through Model.forward -> non_max_suppression on every rank.
  value : device-timed (CUDA events) with the fp32 input batch already resident in HBM.
  e2e   : the same through the public API from HOST memory: pinned uint8 images -> H2D -> model -> NMS ->
          detections copied back to the host (D2H), copies inside the timed region.
N > 1 (torchrun): one process per GPU, batch sharded by rank (weak scaling), ONE NCCL all-gather of the packed
detections per step, issued asynchronously (it overlaps the next step's forward); times are the max over ranks.
--impl reference times the CPU oracle port of the same path on the host cores (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {   # BASELINE.json configs; GFLOP per image = BASELINE.md section 3
    'cfg2': dict(cfg='ablation-ca-scconv-sppfcspc-bifpn.yaml', img=640, bs=64, gflop=162.896),
    'cfg3': dict(cfg='yolov5l-ca-sppfcspc-bifpn-scconv.yaml', img=1536, bs=8, gflop=962.42),
    'cfg4a': dict(cfg='spdconv.yaml', img=1280, bs=32, gflop=816.67),
    'cfg4b': dict(cfg='C3CASPD.yaml', img=1280, bs=32, gflop=800.85),
    'cfg5': dict(cfg=None, img=640, bs=256, gflop=0.0),
}
# SURVEY.md F5 / 8(c) deterministic init: BN running statistics from one train-mode pass over rand(4,3,320,320).
# (A smaller calibration batch leaves the 640x640 network saturated: ~340 k candidates/image all at confidence 1.0.)
CALIB = dict(calib_hw=(320, 320), calib_bs=4)
NMS_KW = dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)
NMS_DETECT_KW = dict(conf_thres=0.25, iou_thres=0.45, multi_label=False, max_det=1000)


def peaks():
    p = dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source='fallback (B200_PROFILING.md)')
    f = ROOT / 'MEASURED_PEAKS.json'
    if f.exists():
        try:
            p.update(json.load(open(f)))
            p['source'] = 'MEASURED_PEAKS.json'
        except Exception:
            pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith('active') for r in self.rows)]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(self.rows))


def cfg5_pred(n, seed=1):
    """rand(n, 25200, 15), xy scaled to 640 px, wh to 4..64 px; obj / cls uniform(0, 1)."""
    import torch
    pred = torch.rand(n, 25200, 15, generator=torch.Generator().manual_seed(seed))
    pred[..., :2] *= 640
    pred[..., 2:4] = pred[..., 2:4] * 60 + 4
    return pred


def cpu_reference_step(cfgd, sd, strides, x):
    """One pass of the CPU oracle port over a batch: forward + decode + NMS (one image per NMS call)."""
    import torch
    from oracle import blocks as O
    from oracle import nms as ON
    with torch.no_grad():
        pred, _, _ = O.forward_model(cfgd, sd, x, strides)
    p = pred.numpy()
    return [ON.non_max_suppression(p[i:i + 1], **NMS_KW)[0] for i in range(p.shape[0])]


def cpu_nms_step(pred_np, kw):
    from oracle import nms as ON
    return [ON.non_max_suppression(pred_np[i:i + 1], **kw)[0] for i in range(pred_np.shape[0])]


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is pure Python
    (no compiled sources to build into oracle/_ref) and /root/reference is absent on the GPU box, so this arm
    times the oracle port (CPU fp32 restatement, pinned against the executed reference) with every host thread."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    import yaml
    from dma_yolo_b200.models import yolo as Y
    from dma_yolo_b200.utils.calib import build_calibrated
    wl = WORKLOADS[args.config]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if args.config == 'cfg5':
        sample_bs = 4
        p = cfg5_pred(sample_bs).numpy()
        step = lambda: cpu_nms_step(p, NMS_KW)
        what = 'non_max_suppression(val-style) on rand(., 25200, 15)'
        sample = f'{args.steps} x {sample_bs} images of 25,200 x 15, numpy NMS port (single thread)'
        cores_used = 1
    else:
        sample_bs = 2 if args.config == 'cfg2' else 1
        m = build_calibrated(wl['cfg'], seed=0, **CALIB)
        sd = {k: v for k, v in m.state_dict().items()}
        cfgd = yaml.safe_load(open(Y.CFG_DIR / wl['cfg']))
        strides = m.stride.tolist()
        x = torch.rand(sample_bs, 3, wl['img'], wl['img'], generator=torch.Generator().manual_seed(1))
        step = lambda: cpu_reference_step(cfgd, sd, strides, x)
        what = f"{wl['cfg']} forward+decode+NMS(val-style) {wl['img']}x{wl['img']}"
        sample = f"{args.steps} x batch {sample_bs} at {wl['img']}x{wl['img']}, torch CPU fp32 + numpy NMS"
        cores_used = torch.get_num_threads()
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = sample_bs * args.steps / dt
    line = dict(metric='images/sec', value=round(v, 3), unit='img/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=round(dt / args.steps * 1e3, 2), higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic', impl='reference',
                config=dict(workload=what, per_step_sample=f"batch {sample_bs} (bounded sample of the batch-{wl['bs']} workload)"),
                cpu_baseline=dict(value=round(v, 3), unit='img/s', cores=cores_used, kind='port', sample=sample),
                e2e=dict(value=round(v, 3), unit='img/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def install_launch_timer(ops, torch):
    """Wrap ops.call so that, while `.on`, every C-ABI launch is bracketed by a CUDA-event pair on the launching stream.
    Kept out of the headline timed region: an event between two launches serialises them and hides the programmatic
    dependent launch overlap."""
    events = []
    real_call = ops.call

    def timed_call(fname, stream, **f):
        if not timed_call.on:
            return real_call(fname, stream, **f)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        real_call(fname, stream, **f)
        e1.record()
        flops, shape, nbytes = 0.0, None, 0.0
        if fname == 'dmay_conv_bn_act':
            flops = 2.0 * f['N'] * f['Ho'] * f['Wo'] * f['Cout'] * f['Cin'] * f['kh'] * f['kw']
            shape = (f['Cin'], f['Cout'], f['kh'], f['stride'], f['Ho'], f['Wo'], int(f.get('residual') is not None),
                     int(f.get('gate_x') is not None))
            # algorithmic bytes of the launch: input + output once, + the residual / gate operand of its epilogue
            opix = float(f['N'] * f['Ho'] * f['Wo'])
            nbytes = (2.0 * f['N'] * f['H'] * f['W'] * f['Cin'] + opix * f['Cout'] * (4 if f.get('out_dtype') == 1 else 2)
                      + 2.0 * opix * f['Cout'] * (int(f.get('residual') is not None) + int(f.get('gate_x') is not None)))
        events.append((fname, e0, e1, flops, shape, nbytes))
    timed_call.on = False
    ops.call = timed_call
    return timed_call, events


def summarize_launches(events, steps):
    """-> (per-entry-point table, conv ms, conv flops, n conv launches)"""
    per = {}
    for fname, e0, e1, *_ in events:
        d = per.setdefault(fname, dict(launches=0, ms=0.0))
        d['launches'] += 1
        d['ms'] += e0.elapsed_time(e1)
    tot = sum(d['ms'] for d in per.values()) or 1.0
    table = {k.replace('dmay_', ''): dict(launches_per_step=d['launches'] // max(steps, 1), ms_per_step=round(d['ms'] / max(steps, 1), 4),
                                          share=round(d['ms'] / tot, 4))
             for k, d in sorted(per.items(), key=lambda kv: -kv[1]['ms'])}
    conv = [e for e in events if e[0] == 'dmay_conv_bn_act']
    return table, sum(e[1].elapsed_time(e[2]) for e in conv), sum(e[3] for e in conv), len(conv)


def run_cfg5(args, torch, dist, D, ops, dev, world, rank, local):
    """NMS stress (BASELINE.json configs[4]): decode-less non_max_suppression on rand(256, 25200, 15), both call styles."""
    from dma_yolo_b200.dist import all_gather_packed_async
    B = args.bs or WORKLOADS['cfg5']['bs']
    host = [cfg5_pred(B, seed=1 + 2 * rank + i).pin_memory() for i in range(2)]
    resident = [h.to(dev) for h in host]
    in_bytes = resident[0].numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def nms(pred, kw):
        return ops.nms_batched(pred, kw['conf_thres'], kw['iou_thres'], multi_label=kw['multi_label'], max_det=kw['max_det'],
                               return_packed=True)

    pending = [None]

    def step(i, kw):
        out, cnt, packed = nms(resident[i & 1], kw)
        if world > 1:
            if pending[0] is not None:
                pending[0].wait()
            pending[0] = all_gather_packed_async(packed, B, kw['max_det'])
        return out, cnt

    def timed(kw):
        for i in range(args.warmup):
            step(i, kw)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            out, cnt = step(i, kw)
        if pending[0] is not None:
            pending[0].wait()
        e1.record()
        barrier()
        return e0.elapsed_time(e1), cnt

    n0 = D.launch_count()
    with ClockSampler(local) as clocks:
        ms_val, cnt_val = timed(NMS_KW)
        launches = D.launch_count() - n0
        ms_det, cnt_det = timed(NMS_DETECT_KW)
        # per-entry-point shares (val style)
        timer, events = install_launch_timer(ops, torch)
        timer.on = True
        for i in range(args.steps):
            step(i, NMS_KW)
        barrier()
        timer.on = False
        table, *_ = summarize_launches(events, args.steps)
        # e2e: host fp32 prediction -> H2D -> NMS -> padded detections D2H
        stage = [torch.empty_like(resident[0]) for _ in range(2)]
        out_host = torch.empty((B, NMS_KW['max_det'], 6), dtype=torch.float32).pin_memory()
        cnt_host = torch.empty((B,), dtype=torch.int32).pin_memory()
        copy_stream = torch.cuda.Stream(dev)
        ready = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        for ev in done:
            ev.record()

        def h2d(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[i & 1])
                stage[i & 1].copy_(host[i & 1], non_blocking=True)
                ready[i & 1].record(copy_stream)

        def e2e_step(i):
            cur = torch.cuda.current_stream()
            cur.wait_event(ready[i & 1])
            dets = D.non_max_suppression(stage[i & 1], **NMS_KW)
            done[i & 1].record(cur)
            out_host.copy_(dets.padded, non_blocking=True)
            cnt_host.copy_(dets.counts, non_blocking=True)
        for i in range(2):
            h2d(i)
            e2e_step(i)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        h2d(0)
        for i in range(args.steps):
            if i + 1 < args.steps:
                h2d(i + 1)
            e2e_step(i)
        t1.record()
        barrier()
        ms_e2e = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms_val, ms_det, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_val, ms_det, ms_e2e = t.tolist()
    pk = peaks()
    per = ms_val / args.steps
    rows_val = B * 25200 * 10            # multi-label expansion: every (row, class) pair is a candidate at 0.001
    achieved = in_bytes / (per / 1e3) / 1e9
    hbm = float(pk.get('hbm_gbs', 6555.0))
    line = dict(
        metric='images/sec', value=round(world * B * args.steps / (ms_val / 1e3), 1), unit='img/s', n_gpus=world, steps=args.steps,
        warmup=args.warmup, ms_per_step=round(per, 3), higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
        data='synthetic',
        config=dict(workload=f'cfg5 NMS stress: non_max_suppression(rand({B},25200,15)) val-style 0.001/0.6/multi_label/300, decode-less, per GPU',
                    l2=f'input {in_bytes >> 20} MB per batch exceeds the 126 MB L2 (two alternating batches); no explicit flush',
                    parallelism=f'dp{world} (batch sharded by rank, one async NCCL all-gather of the packed detections)'),
        candidates_per_s=round(world * rows_val * args.steps / (ms_val / 1e3), 0),
        detect_style=dict(ms_per_step=round(ms_det / args.steps, 3), img_per_s=round(world * B * args.steps / (ms_det / 1e3), 1),
                          nms='0.25/0.45/best-class/1000', mean_detections=float(cnt_det.float().mean())),
        mean_detections=float(cnt_val.float().mean()),
        e2e=dict(value=round(world * B * args.steps / (ms_e2e / 1e3), 1), unit='img/s', h2d_bytes_per_step=world * in_bytes,
                 d2h_bytes_per_step=world * (B * NMS_KW['max_det'] * 6 * 4 + B * 4), ms_per_step=round(ms_e2e / args.steps, 3)),
        gpu_launches=launches, clocks=clocks.summary(), kernels=table,
        roofline=dict(kernel='NMS chain (filter -> top-K select -> sort -> greedy), bytes = the fp32 prediction read once',
                      bound='hbm', achieved=round(achieved, 1), peak=hbm, unit='GB/s', frac=round(achieved / hbm, 4), traffic=None,
                      note='latency / ALU-bound by construction (SURVEY.md 8a a9): read candidates_per_s next to it',
                      peak_source=f"{pk['source']} hbm_gbs"))
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        p = cfg5_pred(4).numpy()
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 10.0 and reps < 6:
            cpu_nms_step(p, NMS_KW)
            reps += 1
        dt = time.perf_counter() - t0
        line['cpu_baseline'] = dict(value=round(4 * reps / dt, 3), unit='img/s', cores=1, kind='port',
                                    sample=f'{reps} x 4 images of 25,200 x 15 (oracle port: numpy NMS with early exit at max_det)')
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='cfg2', choices=sorted(WORKLOADS))
    ap.add_argument('--bs', type=int, default=0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-latency', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import yaml

    import dma_yolo_b200 as D
    from dma_yolo_b200 import ops
    from dma_yolo_b200.dist import all_gather_packed_async
    from dma_yolo_b200.models import yolo as Y
    from dma_yolo_b200.utils.calib import build_calibrated

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback for the product path)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    if args.config == 'cfg5':
        run_cfg5(args, torch, dist, D, ops, dev, world, rank, local)
        if world > 1:
            dist.destroy_process_group()
        return
    wl = WORKLOADS[args.config]
    CFG, IMG = wl['cfg'], wl['img']
    B = args.bs or wl['bs']

    m = build_calibrated(CFG, seed=0, **CALIB)
    sd_cpu = {k: v.clone() for k, v in m.state_dict().items()}
    strides = m.stride.tolist()
    m = m.to(dev).eval()

    g = torch.Generator().manual_seed(100 + rank)
    u8_host = torch.randint(0, 256, (2, B, 3, IMG, IMG), dtype=torch.uint8, generator=g).pin_memory()   # two pinned batches
    x_dev = [(u8_host[i].to(dev).float() / 255) for i in range(2)]                                      # resident fp32 inputs
    max_det = NMS_KW['max_det']
    pending = [None]

    def gather_async(packed):
        # ONE collective per step on NCCL's own stream; the previous step's is waited for first (stream-side), so the
        # exchange of step i overlaps the forward of step i+1 and at most one is in flight
        if pending[0] is not None:
            pending[0].wait()
        pending[0] = all_gather_packed_async(packed, B, max_det)

    def drain():
        if pending[0] is not None:
            pending[0].wait()
            pending[0] = None

    def nms_raw(pred):
        return ops.nms_batched(None, NMS_KW['conf_thres'], NMS_KW['iou_thres'], levels=pred._levels, na=pred._na,
                               nc=pred._no - 5, multi_label=True, max_det=max_det, return_packed=True)

    def step_resident(i):
        with torch.no_grad():
            pred, _ = m(x_dev[i & 1])
            out, cnt, packed = nms_raw(pred)
            if world > 1:
                gather_async(packed)
        return out, cnt

    timer, events = install_launch_timer(ops, torch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for i in range(args.warmup):
        step_resident(i)
    drain()
    barrier()

    # ---- timed: device-resident inputs ----
    n0 = D.launch_count()
    profiling = bool(os.environ.get('DMAY_PROFILE'))   # ncu --profile-from-start off: capture the timed steps only
    with ClockSampler(local) as clocks:
        barrier()
        if profiling:
            torch.cuda.profiler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step_resident(i)
        drain()
        e1.record()
        barrier()
        if profiling:
            torch.cuda.profiler.stop()
        ms = e0.elapsed_time(e1)
        launches = D.launch_count() - n0
        # roofline of the dominant kernel and the per-entry-point shares: the SAME steps again, every launch bracketed by
        # CUDA events on the launching stream (see install_launch_timer for why this is not done in the region above)
        timer.on = True
        for i in range(args.steps):
            step_resident(i)
        drain()
        barrier()
        timer.on = False
        table, conv_ms, conv_flops, n_conv = summarize_launches(events, args.steps)
        if os.environ.get('DMAY_LAYER_TABLE'):
            conv_events = [e for e in events if e[0] == 'dmay_conv_bn_act']
            per = n_conv // max(args.steps, 1)
            rows_ = []
            for i in range(per):
                ts = sorted(conv_events[s_ * per + i][1].elapsed_time(conv_events[s_ * per + i][2]) for s_ in range(args.steps))
                ev = conv_events[i]
                rows_.append(dict(i=i, shape=ev[4], ms=round(ts[len(ts) // 2], 4), tflops=round(ev[3] / ts[len(ts) // 2] / 1e9, 1)))
            json.dump(rows_, open(os.environ['DMAY_LAYER_TABLE'], 'w'))

        # ---- timed: end to end from host memory through the public API ----
        copy_stream = torch.cuda.Stream(dev)
        stage = [torch.empty((B, 3, IMG, IMG), dtype=torch.uint8, device=dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        out_host = torch.empty((B, max_det, 6), dtype=torch.float32).pin_memory()
        cnt_host = torch.empty((B,), dtype=torch.int32).pin_memory()

        done = [torch.cuda.Event() for _ in range(2)]

        def h2d(i):
            # the staging buffer may only be overwritten after the step that last read it has finished
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[i & 1])
                stage[i & 1].copy_(u8_host[i & 1], non_blocking=True)
                ready[i & 1].record(copy_stream)

        def e2e_step(i):
            cur = torch.cuda.current_stream()
            cur.wait_event(ready[i & 1])
            with torch.no_grad():
                pred, _ = m(stage[i & 1])                       # uint8 -> /255 -> SPD bf16 inside the prep kernel
                done[i & 1].record(cur)
                dets = D.non_max_suppression(pred, **NMS_KW)    # public API: list of (n,6) tensors; sizes fetched lazily
            if world > 1:
                gather_async(dets.packed)                       # counts stay on the device until after the exchange
            out_host.copy_(dets.padded, non_blocking=True)      # D2H of the step's result: the batch buffer the list
            cnt_host.copy_(dets.counts, non_blocking=True)      # elements are views of, and the per-image counts
            return dets

        for ev in done:
            ev.record()
        for i in range(2):   # e2e warm-up
            h2d(i)
            e2e_step(i)
        drain()
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        h2d(0)
        for i in range(args.steps):
            if i + 1 < args.steps:
                h2d(i + 1)                                      # next batch's H2D overlaps this batch's compute
            e2e_step(i)
        drain()
        t1.record()
        barrier()
        ms_e2e = t0.elapsed_time(t1)
    clk = clocks.summary()

    # ---- batch-1 latency, detect.py's operating point (detect.py:168-251: bs 1, 0.25 / 0.45 / max_det 1000) ----
    latency = None
    if args.config == 'cfg2' and not args.no_latency:
        try:
            latency = bs1_latency(torch, D, m, dev, IMG)
        except Exception as e:   # never lose the headline line to the auxiliary measurement
            latency = dict(error=f'{type(e).__name__}: {e}'[:200])

    # max over ranks
    ms_rank = ms
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
        allms = [torch.zeros(1, device=dev) for _ in range(world)]
        dist.all_gather(allms, torch.tensor([ms_rank], device=dev))
        per_rank_ms = [round(float(v) / args.steps, 3) for v in allms]
    else:
        per_rank_ms = [round(ms / args.steps, 3)]

    pk = peaks()
    # DRAM traffic of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum summed over the conv launches
    # of ONE step, from the committed ncu launch list of this workload (profiles/<tag>_launches.json, written by
    # tools/summarize_profiles.py; ncu cannot run inside this timed process)
    traffic, traffic_src = None, None
    try:
        if args.config == 'cfg2':
            cands = [f for f in sorted((ROOT / 'profiles').glob('*_launches.json')) if '_cfg' not in f.name]
        else:
            cands = sorted((ROOT / 'profiles').glob(f'*_{args.config}_launches.json'))
        prof = cands[-1]
        pj = json.load(open(prof))
        convs = [v for k, v in pj.items() if 'conv_gemm' in k]
        if convs and B == wl['bs']:
            traffic = float(sum(v['rd'] + v['wr'] for v in convs))
            traffic_src = f'{prof.name}: {sum(v["n"] for v in convs)} conv launches of one step (ncu, cold cache)'
    except Exception:
        pass
    value = world * B * args.steps / (ms / 1e3)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    achieved = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else None
    peak_tf = float(pk.get('bf16_tflops_sustained', 1400.0))
    # what the aggregate could reach: every conv launch at ITS OWN floor max(flops / tensor peak, algorithmic bytes / HBM peak)
    # (the 1x1 layers are HBM-bound: 0.3-0.6 of the tensor peak is their ceiling)
    floor_ms = sum(max(e[3] / (peak_tf * 1e12), e[5] / (float(pk.get('hbm_gbs', 6555.0)) * 1e9)) * 1e3
                   for e in events if e[0] == 'dmay_conv_bn_act')
    in_mb = B * 3 * IMG * IMG * 4 >> 20
    line = dict(
        metric='images/sec', value=round(value, 1), unit='img/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=round(ms / args.steps, 3), higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16',
        data='synthetic',
        config=dict(workload=f'{CFG} forward + Detect decode + NMS(val-style 0.001/0.6/multi_label/300), {IMG}x{IMG}, batch {B}/GPU',
                    weights='random-init seed 0, BN statistics calibrated on rand(4,3,320,320) seed 1 (SURVEY.md F5)',
                    l2=f'inputs ({in_mb} MB fp32/batch, two alternating batches) and activations exceed the 126 MB L2; no explicit flush',
                    parallelism=f'dp{world} (batch sharded by rank, one async NCCL all-gather of the packed detections per step)'),
        e2e=dict(value=round(e2e_value, 1), unit='img/s', h2d_bytes_per_step=world * B * 3 * IMG * IMG,
                 d2h_bytes_per_step=world * (B * max_det * 6 * 4 + B * 4), ms_per_step=round(ms_e2e / args.steps, 3),
                 note='bytes summed over ranks; every rank copies its own slice'),
        gpu_launches=launches, clocks=clk, per_rank_ms_per_step=per_rank_ms,
        roofline=dict(kernel='conv_gemm_kernel (tcgen05 implicit GEMM, all Conv+BN+SiLU layers)', bound='tensor',
                      achieved=round(achieved, 1) if achieved else None, peak=peak_tf, unit='TFLOP/s',
                      frac=round(achieved / peak_tf, 3) if achieved else None, traffic=traffic, traffic_unit='bytes per step (all conv launches)',
                      traffic_source=traffic_src,
                      launches_per_step=n_conv // max(args.steps, 1), share_of_step=round(min(conv_ms / ms_rank, 1.0), 3),
                      measured_over=f'{args.steps} further steps of the same workload right after the timed region, one CUDA-event pair per launch',
                      peak_source=f"{pk['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                      flops_per_step=conv_flops / max(args.steps, 1),
                      ceiling_frac=round(conv_flops / (floor_ms / 1e3) / 1e12 / peak_tf, 3) if floor_ms > 0 else None,
                      frac_of_layer_floors=round(floor_ms / conv_ms, 3) if conv_ms > 0 else None,
                      ceiling_note='ceiling_frac: the aggregate if every launch ran at max(flops / tensor peak, algorithmic bytes / '
                                   'HBM peak); frac_of_layer_floors = sum of those floors / measured conv time'),
        kernels=table,
    )
    if latency is not None:
        line['latency_bs1_ms'] = latency
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded CPU sample of the same workload with the oracle port (a reported baseline, not the target)
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cfgd = yaml.safe_load(open(Y.CFG_DIR / CFG))
        sb = 2 if args.config == 'cfg2' else 1
        xs = torch.rand(sb, 3, IMG, IMG, generator=torch.Generator().manual_seed(1))
        if args.config == 'cfg2':
            cpu_reference_step(cfgd, sd_cpu, strides, xs[:1])
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 12.0 and reps < 8:
            cpu_reference_step(cfgd, sd_cpu, strides, xs)
            reps += 1
        dt = time.perf_counter() - t0
        line['cpu_baseline'] = dict(value=round(sb * reps / dt, 3), unit='img/s', cores=torch.get_num_threads(), kind='port',
                                    sample=f'{reps} x batch {sb} at {IMG}x{IMG} (oracle port: torch CPU fp32 forward + numpy NMS)')
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bs1_latency(torch, D, m, dev, IMG, reps=30):
    """Per-image latency at batch 1 (detect.py's operating point): host-timed from 'image resident on the device' to
    'detections on the host', median over `reps` images; eager launches and, when available, the CUDA-graph replay of
    the same launch sequence (dma_yolo_b200.GraphedDetector)."""
    xs = [torch.rand(1, 3, IMG, IMG, generator=torch.Generator().manual_seed(500 + i)).to(dev) for i in range(4)]
    kw = dict(conf_thres=0.25, iou_thres=0.45, max_det=1000)

    def once(x):
        with torch.no_grad():
            pred, _ = m(x)
            return [d.cpu() for d in D.non_max_suppression(pred, **kw)]
    for i in range(5):
        once(xs[i & 3])
    torch.cuda.synchronize()
    ts = []
    for i in range(reps):
        t0 = time.perf_counter()
        once(xs[i & 3])
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    res = dict(eager=round(ts[len(ts) // 2], 3), nms='detect-style 0.25/0.45/1000',
               note='host wall clock per image, image resident on the device -> detections on the host (NMS and D2H included)',
               v100_reference_ms='6.9 inference + 1.3 NMS (yolov5s, tutorial.ipynb:470-476; other model, other GPU)')
    if hasattr(D, 'GraphedDetector'):
        gd = D.GraphedDetector(m, **kw)
        for i in range(5):
            gd(xs[i & 3])
        torch.cuda.synchronize()
        ts = []
        for i in range(reps):
            t0 = time.perf_counter()
            [d.cpu() for d in gd(xs[i & 3])]
            ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        res['graph'] = round(ts[len(ts) // 2], 3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            gd.replay_only(xs[i & 3])
        e1.record()
        torch.cuda.synchronize()
        res['graph_device_ms'] = round(e0.elapsed_time(e1) / 10, 3)
    return res


if __name__ == '__main__':
    main()
