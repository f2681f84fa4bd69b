"""Import the UNMODIFIED reference from /root/reference (build container only — it does not exist on
the GPU box, so nothing in `-m gpu` tests, smoke() or bench.py may import this module at run time).

Shim: RANK=1 disables the import-time font download in utils/plots.py:64-66;
matplotlib/seaborn are absent and only used for plots -> stub modules; `CA` alias (F3).
"""
import os
import sys
import types

REF = '/root/reference'


def available() -> bool:
    return os.path.isdir(os.path.join(REF, 'models'))


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        m = _Stub(self.__name__ + '.' + k)
        setattr(self, k, m)
        return m

    def __call__(self, *a, **k):
        return self


_loaded = None


def load():
    """-> namespace with Model, non_max_suppression, common (module), yolo (module), val (module or None)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('reference tree not present')
    os.environ.setdefault('RANK', '1')
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.font_manager', 'seaborn'):
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
    # our package may have registered aliases under the same names; the reference must win here
    for name in list(sys.modules):
        if name == 'models' or name.startswith('models.') or name == 'utils' or name.startswith('utils.'):
            mod = sys.modules[name]
            if not getattr(mod, '__file__', '') or not str(getattr(mod, '__file__', '')).startswith(REF):
                del sys.modules[name]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import models.common as C
    import models.yolo as Y
    from utils.general import non_max_suppression
    Y.CA = C.CoorAttention
    ns = types.SimpleNamespace(Model=Y.Model, non_max_suppression=non_max_suppression, common=C, yolo=Y)
    _loaded = ns
    return ns
