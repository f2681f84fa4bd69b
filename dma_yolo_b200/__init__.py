"""dma_yolo_b200 — B200-native (sm_100a) detection-forward hot path of DMA-YOLO.

    from dma_yolo_b200 import Model, non_max_suppression
    model = Model('ablation-ca-scconv-sppfcspc-bifpn.yaml').cuda().eval()
    pred, _ = model(img)                                  # kernels: conv/BN/SiLU, SCConv, CoordAtt, SPPFCSPC, BiFPN
    dets = non_max_suppression(pred, 0.001, 0.6, multi_label=True)   # fused decode + filter + batched NMS

`install_aliases()` additionally registers the package's `models.*` / `utils.*` modules under the
reference's import paths (`models.yolo`, `models.common`, `models.cspcm`, `utils.general`), which is
what pickled DMA-YOLO checkpoints and detect.py / val.py import.
(The directory is also reachable as `dma-yolo_b200/`; a hyphen cannot appear in a Python identifier.)
"""
from __future__ import annotations

import sys

from . import _lib, ops  # noqa: F401
from ._lib import DmayError, launch_count  # noqa: F401
from .graph import GraphedDetector  # noqa: F401
from .models.common import set_backend  # noqa: F401
from .models.yolo import Detect, Model, parse_model  # noqa: F401
from .utils.general import non_max_suppression, scale_coords, xywh2xyxy, xyxy2xywh  # noqa: F401

__version__ = "0.1.0"


def install_aliases(force: bool = False):
    """Expose this package as the reference's top-level `models` / `utils` packages."""
    from . import models, utils
    from .models import common, cspcm, detect_t, yolo
    from .utils import augmentations, general, metrics, torch_utils
    table = {"models": models, "models.common": common, "models.cspcm": cspcm, "models.yolo": yolo,
             "models.detect_t": detect_t,
             "utils": utils, "utils.general": general, "utils.torch_utils": torch_utils,
             "utils.augmentations": augmentations, "utils.metrics": metrics}
    for name, mod in table.items():
        if force or name not in sys.modules:
            sys.modules[name] = mod
