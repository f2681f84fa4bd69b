"""Shared helpers for the test-suite."""
from pathlib import Path

import numpy as np
import torch

GOLD = Path(__file__).resolve().parent / 'golden'


def load_golden(name):
    z = np.load(GOLD / f'{name}.npz', allow_pickle=False)
    d = {k: (torch.from_numpy(z[k]) if z[k].dtype.kind in 'fiub' else z[k]) for k in z.files}
    sd = {k[3:]: v for k, v in d.items() if k.startswith('sd/')}
    ins = [d[f'in{i}'] for i in range(8) if f'in{i}' in d]
    return d, sd, ins


def assert_close(a, b, atol, rtol, what=''):
    a, b = a.float().cpu(), b.float().cpu()
    assert a.shape == b.shape, f'{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}'
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bad.any(), (f'{what}: {int(bad.sum())}/{bad.numel()} elements out of tolerance; '
                           f'max abs err {float(err.max()):.4g} at ref {float(b.flatten()[err.argmax()]):.4g}')
