"""Re-emit reference model configs (pure data: nc / multiples / anchors / layer table) in a normalised
form under dma_yolo_b200/models/, so the drop-in `Model('<name>.yaml')` works where /root/reference
does not exist (GPU box).  Run in the build container:  python tools/import_cfgs.py [names...]

Only the layer table is carried over (parsed with yaml.safe_load and re-serialised, one layer per
line); comments and formatting of the source files are not.
"""
import sys
from pathlib import Path

import yaml

SRC = Path('/root/reference/models')
DST = Path(__file__).resolve().parent.parent / 'dma_yolo_b200' / 'models'
DEFAULT = ['yolov5s', 'yolov5n', 'yolov5m', 'yolov5l', 'yolov5x', 'ablation-ca-scconv-sppfcspc-bifpn',
           'yolov5l-ca-sppfcspc-bifpn-scconv', 'spdconv', 'C3CASPD', 'CASPD_ODRTA']


def all_names():
    """Every reference config with a layer table (models/*.yaml and models/hub/*.yaml; hub/anchors.yaml is a list of anchors)."""
    out = []
    for f in sorted(SRC.glob('*.yaml')) + sorted((SRC / 'hub').glob('*.yaml')):
        d = yaml.safe_load(f.read_text(errors='ignore'))
        if isinstance(d, dict) and 'backbone' in d and 'anchors' in d:
            out.append(str(f.relative_to(SRC))[:-5])
    return out


def emit(name: str):
    d = yaml.safe_load((SRC / f'{name}.yaml').read_text(errors='ignore'))
    (DST / f'{name}.yaml').parent.mkdir(exist_ok=True)
    flow = lambda v: yaml.safe_dump(v, default_flow_style=True, width=10 ** 6).strip()
    lines = [f'# {name}: layer table for dma_yolo_b200.Model (normalised by tools/import_cfgs.py)',
             f"nc: {d['nc']}", f"depth_multiple: {d['depth_multiple']}", f"width_multiple: {d['width_multiple']}"]
    if isinstance(d['anchors'], list):
        lines.append('anchors:')
        lines += [f'  - {flow(a)}' for a in d['anchors']]
    else:
        lines.append(f"anchors: {d['anchors']}")
    for sec in ('backbone', 'head'):
        lines.append(f'{sec}:')
        lines += [f'  - {flow(layer)}' for layer in d[sec]]
    (DST / f'{name}.yaml').write_text('\n'.join(lines) + '\n')
    back = yaml.safe_load((DST / f'{name}.yaml').read_text())
    assert back == d, name
    print('wrote', name)


if __name__ == '__main__':
    names = sys.argv[1:] or DEFAULT
    if names == ['--all']:
        names = all_names()
    for n in names:
        emit(n)
