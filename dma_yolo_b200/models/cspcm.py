"""models/cspcm.py mirror: the reference imports `*` from here AFTER models.common inside models/yolo.py,
so top-level YAML `Conv` layers (and pickled checkpoints) resolve to `models.cspcm.Conv`
(models/cspcm.py:11-23).  Same arithmetic, same kernels."""
from .common import Conv as _CommonConv
from .common import autopad  # noqa: F401


class Conv(_CommonConv):
    """Standard convolution — models/cspcm.py:11-23 (byte-equivalent duplicate of models/common.py:50-77)."""
