"""Deferred Detect output.

`Detect.forward` must return the dense `(bs, rows, no)` prediction (models/yolo.py:103), but the
confidence threshold that lets decode be fused with the candidate filter is only known to
`non_max_suppression` (utils/general.py:633).  `LazyPred` is a real `torch.Tensor` subclass with the
dense tensor's shape/dtype/device that carries the raw head logits; our `non_max_suppression`
consumes the logits directly (fused decode + filter, nothing dense is written), and ANY other torch
operation on it transparently materialises the dense tensor with the decode kernel first.
"""
from __future__ import annotations

import torch
from torch.utils._pytree import tree_map

from . import ops


class LazyPred(torch.Tensor):
    @staticmethod
    def __new__(cls, levels, na, no):
        lg = levels[0].logits
        rows = sum(na * lv.ny * lv.nx for lv in levels)
        r = torch.Tensor._make_wrapper_subclass(cls, (lg.shape[0], rows, no), dtype=torch.float32, device=lg.device,
                                                requires_grad=False)
        r._levels, r._na, r._no, r._dense = levels, na, no, None
        return r

    def __init__(self, levels, na, no):
        pass

    def dense(self) -> torch.Tensor:
        if self._dense is None:
            self._dense = ops.detect_decode(self._levels, self._na, self._no)
        return self._dense

    def __repr__(self):
        return f"LazyPred(shape={tuple(self.shape)}, materialised={self._dense is not None})"

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        un = lambda t: t.dense() if isinstance(t, LazyPred) else t
        return func(*tree_map(un, args), **tree_map(un, kwargs or {}))
