"""Headline benchmark: images/s of the detection forward hot path (forward + Detect decode + NMS).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): ablation-ca-scconv-sppfcspc-bifpn.yaml, 640x640, batch 64 per GPU,
bf16 activations, synthetic images, seeded BN-calibrated random-init weights, val-style NMS
(conf 0.001, iou 0.6, multi_label, max_det 300 — val.py:235).

One "step" = one batch through Model.forward -> non_max_suppression on every rank.
  value : device-timed (CUDA events) with the fp32 input batch already resident in HBM.
  e2e   : the same through the public API from HOST memory: pinned uint8 images -> H2D -> model -> NMS ->
          detections copied back to the host (D2H), copies inside the timed region.
N > 1 (torchrun): one process per GPU, batch sharded by rank (weak scaling), NCCL all-gather of the padded
detections each step; times are the max over ranks.  --impl reference times the CPU oracle port of the same
path on the host cores (rank 0 only).
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

CFG = 'ablation-ca-scconv-sppfcspc-bifpn.yaml'
IMG, BS = 640, 64
# SURVEY.md F5 / 8(c) deterministic init: BN running statistics from one train-mode pass over rand(4,3,320,320).
# (A smaller calibration batch leaves the 640x640 network saturated: ~340 k candidates/image all at confidence 1.0.)
CALIB = dict(calib_hw=(320, 320), calib_bs=4)
NMS_KW = dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)
GFLOP_PER_IMG = 162.896  # conv/linear FLOPs of cfg-2 per image


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


def cpu_reference_step(cfgd, sd, strides, x):
    """One pass of the CPU oracle port over a batch: forward + decode + NMS (one image per NMS call)."""
    import torch
    from oracle import blocks as O
    from oracle import nms as ON
    with torch.no_grad():
        pred, _, _ = O.forward_model(cfgd, sd, x, strides)
    p = pred.numpy()
    return [ON.non_max_suppression(p[i:i + 1], **NMS_KW)[0] for i in range(p.shape[0])]


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
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_bs = 2
    m = build_calibrated(CFG, seed=0, **CALIB)
    sd = {k: v for k, v in m.state_dict().items()}
    cfgd = yaml.safe_load(open(Y.CFG_DIR / CFG))
    strides = m.stride.tolist()
    x = torch.rand(sample_bs, 3, IMG, IMG, generator=torch.Generator().manual_seed(1))
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_reference_step(cfgd, sd, strides, x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(cfgd, sd, strides, x)
    dt = time.perf_counter() - t0
    v = sample_bs * args.steps / dt
    line = dict(metric='images/sec', value=round(v, 3), unit='img/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=round(dt / args.steps * 1e3, 2), higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic', impl='reference',
                config=dict(workload=f'{CFG} forward+decode+NMS(val-style) 640x640', per_step_sample=f'batch {sample_bs} (bounded sample of the batch-{BS} workload)'),
                cpu_baseline=dict(value=round(v, 3), unit='img/s', cores=torch.get_num_threads(), kind='port',
                                  sample=f'{args.steps} x batch {sample_bs} at 640x640, torch CPU fp32 + numpy NMS'),
                e2e=dict(value=round(v, 3), unit='img/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--bs', type=int, default=BS)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import yaml

    import dma_yolo_b200 as D
    from dma_yolo_b200 import ops
    from dma_yolo_b200.dist import all_gather_detections
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
    B = args.bs

    m = build_calibrated(CFG, seed=0, **CALIB)
    sd_cpu = {k: v.clone() for k, v in m.state_dict().items()}
    strides = m.stride.tolist()
    m = m.to(dev).eval()

    g = torch.Generator().manual_seed(100 + rank)
    u8_host = torch.randint(0, 256, (2, B, 3, IMG, IMG), dtype=torch.uint8, generator=g).pin_memory()   # two pinned batches
    x_dev = [(u8_host[i].to(dev).float() / 255) for i in range(2)]                                      # resident fp32 inputs
    max_det = NMS_KW['max_det']

    def step_resident(i):
        with torch.no_grad():
            pred, _ = m(x_dev[i & 1])
            out, cnt = nms_raw(pred)
            if world > 1:
                all_gather_detections(out, cnt)
        return out, cnt

    def nms_raw(pred):
        return ops.nms_batched(None, NMS_KW['conf_thres'], NMS_KW['iou_thres'], levels=pred._levels, na=pred._na,
                               nc=pred._no - 5, multi_label=True, max_det=max_det)

    # ---- conv launch timing hook (roofline of the dominant kernel, measured live in the timed region) ----
    conv_events = []
    real_call = ops.call

    def timed_call(fname, stream, **f):
        if fname == 'dmay_conv_bn_act' and timed_call.on:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            real_call(fname, stream, **f)
            e1.record()
            flops = 2.0 * f['N'] * f['Ho'] * f['Wo'] * f['Cout'] * f['Cin'] * f['kh'] * f['kw']
            conv_events.append((e0, e1, flops, (f['Cin'], f['Cout'], f['kh'], f['stride'], f['Ho'], f['Wo'],
                                                 int(f.get('residual') is not None), int(f.get('gate_x') is not None))))
        else:
            real_call(fname, stream, **f)
    timed_call.on = False
    ops.call = timed_call

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for i in range(args.warmup):
        step_resident(i)
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
        e1.record()
        barrier()
        if profiling:
            torch.cuda.profiler.stop()
        ms = e0.elapsed_time(e1)
        launches = D.launch_count() - n0
        # roofline of the dominant kernel: the SAME steps again, every conv launch bracketed by CUDA events on the
        # launching stream.  Kept out of the region above because an event between two launches serialises them
        # (the convs use programmatic dependent launch to overlap their preamble with the previous kernel's tail).
        timed_call.on = True
        for i in range(args.steps):
            step_resident(i)
        barrier()
        timed_call.on = False
        conv_ms = sum(a.elapsed_time(b) for a, b, *_ in conv_events)
        conv_flops = sum(ev[2] for ev in conv_events)
        n_conv = len(conv_events)
        if os.environ.get('DMAY_LAYER_TABLE'):
            per = n_conv // max(args.steps, 1)
            rows_ = []
            for i in range(per):
                ts = sorted(conv_events[s_ * per + i][0].elapsed_time(conv_events[s_ * per + i][1]) for s_ in range(args.steps))
                ev = conv_events[i]
                rows_.append(dict(i=i, shape=ev[3], ms=round(ts[len(ts) // 2], 4), tflops=round(ev[2] / ts[len(ts) // 2] / 1e9, 1)))
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
                dets = D.non_max_suppression(pred, **NMS_KW)    # public API: list of (n,6) tensors (syncs on counts)
            if world > 1:
                all_gather_detections(dets.padded, dets.counts)
            out_host.copy_(dets.padded, non_blocking=True)      # D2H of the step's result: the batch buffer the list
            cnt_host.copy_(dets.counts, non_blocking=True)      # elements are views of, and the per-image counts
            return dets

        for ev in done:
            ev.record()
        for i in range(2):   # e2e warm-up
            h2d(i)
            e2e_step(i)
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        h2d(0)
        for i in range(args.steps):
            if i + 1 < args.steps:
                h2d(i + 1)                                      # next batch's H2D overlaps this batch's compute
            e2e_step(i)
        t1.record()
        barrier()
        ms_e2e = t0.elapsed_time(t1)
    clk = clocks.summary()

    # max over ranks
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())

    pk = peaks()
    # DRAM traffic of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum summed over the conv launches
    # of ONE step, from the committed ncu launch list of this workload (profiles/<tag>_launches.json, written by
    # tools/summarize_profiles.py; ncu cannot run inside this timed process)
    traffic, traffic_src = None, None
    try:
        prof = sorted((ROOT / 'profiles').glob('*_launches.json'))[-1]
        pj = json.load(open(prof))
        convs = [v for k, v in pj.items() if 'conv_gemm' in k]
        if convs and B == BS:
            traffic = float(sum(v['rd'] + v['wr'] for v in convs))
            traffic_src = f'{prof.name}: {sum(v["n"] for v in convs)} conv launches of one step (ncu, cold cache)'
    except Exception:
        pass
    value = world * B * args.steps / (ms / 1e3)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    achieved = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else None
    peak_tf = float(pk.get('bf16_tflops_sustained', 1400.0))
    line = dict(
        metric='images/sec', value=round(value, 1), unit='img/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=round(ms / args.steps, 3), higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16',
        data='synthetic',
        config=dict(workload=f'{CFG} forward + Detect decode + NMS(val-style 0.001/0.6/multi_label/300), 640x640, batch {B}/GPU',
                    weights='random-init seed 0, BN statistics calibrated on rand(4,3,320,320) seed 1 (SURVEY.md F5)', l2='inputs (315 MB fp32/batch) and activations exceed the 126 MB L2; no explicit flush',
                    parallelism=f'dp{world} (batch sharded by rank, NCCL all-gather of detections)'),
        e2e=dict(value=round(e2e_value, 1), unit='img/s', h2d_bytes_per_step=world * B * 3 * IMG * IMG,
                 d2h_bytes_per_step=world * (B * max_det * 6 * 4 + B * 4), ms_per_step=round(ms_e2e / args.steps, 3),
                 note='bytes summed over ranks; every rank copies its own slice'),
        gpu_launches=launches, clocks=clk,
        roofline=dict(kernel='conv_gemm_kernel (tcgen05 implicit GEMM, all Conv+BN+SiLU layers)', bound='tensor',
                      achieved=round(achieved, 1) if achieved else None, peak=peak_tf, unit='TFLOP/s',
                      frac=round(achieved / peak_tf, 3) if achieved else None, traffic=traffic, traffic_unit='bytes per step (all conv launches)',
                      traffic_source=traffic_src,
                      launches_per_step=n_conv // max(args.steps, 1), share_of_step=round(min(conv_ms / ms, 1.0), 3),
                      measured_over=f'{args.steps} further steps of the same workload right after the timed region, one CUDA-event pair per launch',
                      peak_source=f"{pk['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                      flops_per_step=conv_flops / max(args.steps, 1)),
    )
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded CPU sample of the same workload with the oracle port (a reported baseline, not the target)
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cfgd = yaml.safe_load(open(Y.CFG_DIR / CFG))
        xs = torch.rand(2, 3, IMG, IMG, generator=torch.Generator().manual_seed(1))
        cpu_reference_step(cfgd, sd_cpu, strides, xs[:1])
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 12.0 and reps < 8:
            cpu_reference_step(cfgd, sd_cpu, strides, xs)
            reps += 1
        dt = time.perf_counter() - t0
        line['cpu_baseline'] = dict(value=round(2 * reps / dt, 3), unit='img/s', cores=torch.get_num_threads(), kind='port',
                                    sample=f'{reps} x batch 2 at 640x640 (oracle port: torch CPU fp32 forward + numpy NMS)')
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
