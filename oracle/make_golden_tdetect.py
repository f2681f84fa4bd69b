"""Generate tests/golden/tdetect_*.npz by EXECUTING the reference's TDetect head (build container only).

    python -m oracle.make_golden_tdetect
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent


def main():
    sys.path.insert(0, str(ROOT))
    from oracle import refshim
    from oracle.make_golden import randomize_bn, rnd, round_module, save
    refshim.load()
    from models.detect_t import TDetect      # the reference's module (refshim put /root/reference on sys.path)
    g = torch.Generator().manual_seed(91)
    torch.manual_seed(91)

    def fixture(name, nc, ch, sizes, strides, seed):
        gg = torch.Generator().manual_seed(seed)
        mod = TDetect(nc, ch)
        mod.stride = torch.tensor(strides, dtype=torch.float32)
        mod.bias_init()
        randomize_bn(mod, gg)
        with torch.no_grad():
            for n_, p_ in mod.named_parameters():
                if n_.endswith('2.bias'):
                    p_.add_(torch.randn(p_.shape, generator=gg) * 0.5)
        round_module(mod)
        mod.eval()
        xs = [rnd(g, 2, c, h, w) for c, (h, w) in zip(ch, sizes)]
        with torch.no_grad():
            y, (feats, box, cls) = mod([x.clone() for x in xs])
        arrs = {f'in{i}': x for i, x in enumerate(xs)}
        arrs.update(out=y, box=box, cls=cls, strides=torch.tensor(strides, dtype=torch.float32))
        for k, v in mod.state_dict().items():
            arrs['sd/' + k] = v
        save(name, **arrs)

    fixture('tdetect_nc10_3lv', 10, (64, 128, 256), [(12, 10), (6, 5), (3, 3)], [8., 16., 32.], 1)
    fixture('tdetect_nc20_4lv', 20, (32, 64, 64, 128), [(16, 12), (8, 6), (4, 3), (2, 2)], [4., 8., 16., 32.], 2)


if __name__ == '__main__':
    main()
