"""tests/golden/yaml_digests.json: for EVERY reference config (models/*.yaml, models/hub/*.yaml), does the UNMODIFIED
reference build it (seed 0), and if so the sha256 digest of its state_dict and its parameter count.  The product's
`Model(<same yaml>)` must build exactly the same set with identical digests (tests/test_modules_cpu.py).

    python -m oracle.make_golden_yaml_digests          (build container only)
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    from dma_yolo_b200.utils.calib import state_digest
    from oracle import refshim
    R = refshim.load()
    src = Path('/root/reference/models')
    out = {}
    for f in sorted(src.glob('*.yaml')) + sorted((src / 'hub').glob('*.yaml')):
        name = str(f.relative_to(src))[:-5]
        try:
            torch.manual_seed(0)
            m = R.Model(str(f))
            out[name] = dict(ok=True, digest=state_digest(m.state_dict()), params=sum(p.numel() for p in m.parameters()),
                             keys=len(m.state_dict()))
        except Exception as e:   # the reference itself cannot build this file (missing args, wrong channel counts ...)
            out[name] = dict(ok=False, error=type(e).__name__)
        print(name, out[name], flush=True)
    json.dump(out, open(ROOT / 'tests' / 'golden' / 'yaml_digests.json', 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
