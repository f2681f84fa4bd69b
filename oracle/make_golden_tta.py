"""tests/golden/tta_conditioned.npz: `model(x, augment=True)` (models/yolo.py:194-209) of the UNMODIFIED reference on the
conditioned checkpoint (tests/golden/conditioned_ablation.pt) and two synthetic images of 256 x 320 (build container only).

    python -m oracle.make_golden_tta
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    from oracle import refshim, synth
    R = refshim.load()
    ck = torch.load(ROOT / 'tests' / 'golden' / 'conditioned_ablation.pt', map_location='cpu')
    m = R.Model(ck['cfg'])
    m.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in ck['state_dict'].items()})
    m.eval()
    im, _ = synth.make_batch(4242, 2, 320)
    x = (torch.from_numpy(im[:, :, :256, :]).float() / 255).bfloat16().float().contiguous()      # non-square: 256 x 320
    with torch.no_grad():
        out = m(x, augment=True)[0]
        plain = m(x)[0]
    np.savez_compressed(ROOT / 'tests' / 'golden' / 'tta_conditioned.npz', x=x.numpy(), out=out.numpy(), plain=plain.numpy())
    print('tta_conditioned.npz', tuple(out.shape), tuple(plain.shape))


if __name__ == '__main__':
    main()
