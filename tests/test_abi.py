"""The C-ABI library loads and exports every symbol include/dmayolo.h declares (no compute calls)."""
import ctypes
import subprocess

import pytest

from dma_yolo_b200 import _lib


def test_header_parses():
    structs, funcs, consts = _lib.parse_header()
    assert 'dmay_conv_params' in structs and 'dmay_nms_params' in structs
    assert funcs['dmay_conv_bn_act'][0] == 'int'
    assert consts['DMAY_EUNSUPPORTED'] == -2


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    for name in _lib.exported_symbols():
        assert hasattr(L, name), name
    assert L.dmay_version() >= 100
    assert isinstance(_lib.launch_count(), int)
    assert 'unsupported' in _lib.strerror(-2)


def test_struct_layout_matches_compiler():
    """sizeof() of every params struct as seen by gcc equals the ctypes view (no padding surprises)."""
    names = sorted(_lib.STRUCTS)
    src = '#include <stdio.h>\n#include "dmayolo.h"\nint main(){' + ''.join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + 'return 0;}'
    import tempfile, os
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, 't.c')
        open(c, 'w').write(src)
        exe = os.path.join(td, 't')
        subprocess.run(['gcc', '-I', str(_lib.HEADER.parent), c, '-o', exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    sizes = dict(line.split() for line in out.strip().splitlines())
    for n in names:
        assert int(sizes[n]) == ctypes.sizeof(_lib.STRUCTS[n]), n


def test_null_params_rejected_without_gpu():
    """Argument validation happens before any CUDA call: a zeroed struct returns DMAY_EINVAL."""
    L = _lib.lib()
    for fname in ('dmay_spd', 'dmay_adconcat', 'dmay_scconv_gate', 'dmay_sppf_pool3', 'dmay_coordatt',
                  'dmay_conv_bn_act', 'dmay_nms_greedy', 'dmay_detect_decode'):
        sname = _lib._FUNCS[fname][1][0].split()[1].rstrip('*')
        st = _lib.STRUCTS[sname]()
        assert getattr(L, fname)(ctypes.byref(st), None) == _lib.CONSTS['DMAY_EINVAL'], fname
    with pytest.raises(AttributeError):
        _lib.call('dmay_spd', 0, not_a_field=1)
