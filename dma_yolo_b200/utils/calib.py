"""Deterministic, non-vacuous weights for tests and benchmarks.

Random-init BN running stats (0/1) make activations decay ~10x per stage, so every tolerance passes
trivially.  Recipe: seed -> build -> every BatchNorm momentum=1 -> one train-mode forward on a seeded
batch (running stats := batch stats) -> eval -> round floats to bf16 (so fp32 oracles and the bf16
kernels hold identical weights).  oracle/make_golden.py applies the same recipe to the reference Model.
"""
from __future__ import annotations

import hashlib

import torch
import torch.nn as nn


def round_module_bf16(mod: nn.Module):
    for p in mod.parameters():
        p.data = p.data.bfloat16().float()
    for _, b in mod.named_buffers():
        if b.is_floating_point():
            b.data = b.data.bfloat16().float()


def state_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        v = sd[k]
        if torch.is_tensor(v):
            h.update(k.encode())
            h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build_calibrated(cfg, seed=0, nc=None, calib_hw=(128, 128), calib_bs=2, model_cls=None):
    if model_cls is None:
        from ..models.yolo import Model as model_cls
    torch.manual_seed(seed)
    m = model_cls(cfg, nc=nc) if nc else model_cls(cfg)
    for mod in m.modules():
        if isinstance(mod, nn.BatchNorm2d):
            mod.momentum = 1.0
    m.train()
    with torch.no_grad():
        m(torch.rand(calib_bs, 3, *calib_hw, generator=torch.Generator().manual_seed(seed + 1)))
    m.eval()
    for mod in m.modules():
        if isinstance(mod, nn.BatchNorm2d):
            mod.momentum = 0.03
    round_module_bf16(m)
    return m
