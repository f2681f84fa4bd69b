"""Generate tests/golden/swin_*.npz by EXECUTING the reference's SwinTransformerLayer / C3STR (build container only).

    python -m oracle.make_golden_swin
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
    R = refshim.load()
    C = R.common
    g = torch.Generator().manual_seed(77)
    torch.manual_seed(77)

    def fixture(name, mod, x, seed):
        gg = torch.Generator().manual_seed(seed)
        randomize_bn(mod, gg)
        with torch.no_grad():
            for n_, p_ in mod.named_parameters():       # non-trivial LayerNorm affine / biases / position table
                if 'norm' in n_ or n_.endswith('bias') or 'relative_position_bias_table' in n_:
                    p_.add_(torch.randn(p_.shape, generator=gg) * 0.2)
        round_module(mod)
        mod.eval()
        with torch.no_grad():
            out = mod(x.clone())
        arrs = {'in0': x, 'out': out}
        for k, v in mod.state_dict().items():
            arrs['sd/' + k] = v
        save(name, **arrs)

    fixture('swin_layer_16x24', C.SwinTransformerLayer(64, num_heads=2, window_size=8, shift_size=0), rnd(g, 2, 64, 16, 24), 1)
    fixture('swin_layer_shift_16x24', C.SwinTransformerLayer(64, num_heads=2, window_size=8, shift_size=4), rnd(g, 2, 64, 16, 24), 2)
    fixture('swin_layer_shift_13x10', C.SwinTransformerLayer(64, num_heads=2, window_size=8, shift_size=4), rnd(g, 1, 64, 13, 10), 3)
    fixture('swin_layer_13x10', C.SwinTransformerLayer(32, num_heads=1, window_size=8, shift_size=0), rnd(g, 1, 32, 13, 10), 4)
    fixture('c3str_n2_20x12', C.C3STR(64, 128, 2, False), rnd(g, 2, 64, 20, 12), 5)
    # HorNet blocks: gamma1 / gamma2 default to 1e-6 (the block would be an identity): randomised to O(1)
    def hor(name, mod, x, seed):
        gg = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for n_, p_ in mod.named_parameters():
                if 'gamma' in n_:
                    p_.copy_(torch.randn(p_.shape, generator=gg) * 0.5)
        fixture(name, mod, x, seed)
    hor('horblock_64_12x10', C.HorBlock(64), rnd(g, 2, 64, 12, 10), 6)
    hor('horblock_128_9x7', C.HorBlock(128), rnd(g, 1, 128, 9, 7), 7)
    hor('c3hb_n2_16x12', C.C3HB(64, 128, 2, False), rnd(g, 2, 64, 16, 12), 8)


if __name__ == '__main__':
    main()
