"""Drop-in boundary on CPU: same class names / ctor signatures / state_dict keys as the reference, same
results from the torch-op bodies, same seeded weights from `Model(cfg)` (RNG consumption order)."""
import numpy as np
import pytest
import torch
import yaml

import dma_yolo_b200 as D
from dma_yolo_b200.models import common as C
from dma_yolo_b200.models import yolo as Y
from dma_yolo_b200.utils.calib import build_calibrated, state_digest
from oracle import blocks as O
from tests.util import assert_close, load_golden

CTORS = {
    'conv_k3s1': lambda: C.Conv(16, 32, 3, 1), 'conv_k3s2_odd': lambda: C.Conv(16, 32, 3, 2),
    'conv_k1': lambda: C.Conv(32, 16, 1, 1), 'conv_stem_k6s2p2': lambda: C.Conv(3, 32, 6, 2, 2),
    'conv_c64_k3': lambda: C.Conv(64, 64, 3, 1), 'bottleneck': lambda: C.Bottleneck(32, 32, True, e=1.0),
    'c3_n2': lambda: C.C3(32, 32, 2), 'c3_n1_noshortcut': lambda: C.C3(64, 32, 1, False),
    'coordatt_7x5': lambda: C.CoorAttention(64, 64), 'coordatt_20x20': lambda: C.CA(256, 256),
    'spd': lambda: C.space_to_depth(), 'scconv_38': lambda: C.SCConv(16, 32, 2),
    'scconv_16x12': lambda: C.SCConv(32, 64, 2), 'scconv_19x23_s1': lambda: C.SCConv(16, 16, 1),
    'adconcat2': lambda: C.AdConcat2(), 'adconcat3': lambda: C.AdConcat3(), 'adapt_add2': lambda: C.Adapt_Add2(),
    'adapt_add3': lambda: C.Adapt_Add3(16, 16, 32), 'sppf_20': lambda: C.SPPF(32, 32, 5),
    'sppfcspc_12x9': lambda: C.SPPFCSPC(32, 32), 'spp': lambda: C.SPP(32, 32), 'sppcspc': lambda: C.SPPCSPC(32, 32),
    'swin_layer_16x24': lambda: C.SwinTransformerLayer(64, num_heads=2, window_size=8, shift_size=0),
    'swin_layer_shift_16x24': lambda: C.SwinTransformerLayer(64, num_heads=2, window_size=8, shift_size=4),
    'swin_layer_shift_13x10': lambda: C.SwinTransformerLayer(64, num_heads=2, window_size=8, shift_size=4),
    'swin_layer_13x10': lambda: C.SwinTransformerLayer(32, num_heads=1, window_size=8, shift_size=0),
    'c3str_n2_20x12': lambda: C.C3STR(64, 128, 2, False),
    'horblock_64_12x10': lambda: C.HorBlock(64), 'horblock_128_9x7': lambda: C.HorBlock(128),
    'c3hb_n2_16x12': lambda: C.C3HB(64, 128, 2, False),
}


def make_module(name):
    d, sd, ins = load_golden(name)
    m = CTORS[name]()
    if sd:
        m.load_state_dict(sd, strict=True)
    for b in m.modules():
        if isinstance(b, torch.nn.BatchNorm2d):
            b.eps = 1e-3
    return m.eval(), d, ins


@pytest.mark.parametrize('name', list(CTORS))
def test_cpu_body_matches_reference(name):
    m, d, ins = make_module(name)
    with torch.no_grad():
        y = m([t.clone() for t in ins] if len(ins) > 1 else ins[0].clone())
    assert_close(y, d['out'], atol=2e-5, rtol=2e-5, what=name)


def test_detect_cpu_and_keys():
    d, sd, ins = load_golden('detect_nc4')
    det = Y.Detect(nc=4, anchors=[[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]],
                   ch=(16, 32, 64))
    det.stride = torch.tensor([8., 16., 32.])
    det.load_state_dict(sd, strict=True)
    det.eval()
    with torch.no_grad():
        pred, raw = det([t.clone() for t in ins])
    assert_close(pred, d['out'], atol=1e-4, rtol=1e-5, what='detect')


@pytest.mark.parametrize('cfg', ['yolov5s', 'ablation-ca-scconv-sppfcspc-bifpn'])
def test_model_seeded_weights_and_forward(cfg):
    d, _, ins = load_golden('model_' + cfg)
    m = build_calibrated(cfg + '.yaml', seed=0)
    assert state_digest(m.state_dict()) == str(d['digest']), 'seeded weights differ from the reference Model'
    assert torch.equal(m.stride, d['strides'])
    with torch.no_grad():
        pred, raw = m(ins[0])
    assert_close(pred, d['out'], atol=1e-4, rtol=1e-4, what='pred')
    # oracle interpreter on the same weights
    cfgd = yaml.safe_load(open(Y.CFG_DIR / (cfg + '.yaml')))
    with torch.no_grad():
        po, _, _ = O.forward_model(cfgd, m.state_dict(), ins[0], m.stride.tolist())
    assert_close(po, d['out'], atol=1e-3, rtol=2e-4, what='oracle pred')  # fp32 re-association over ~100 convs
    # CPU NMS body == reference outputs
    for style, kw in (('detect', dict(conf_thres=0.25, iou_thres=0.45, max_det=1000)),
                      ('val', dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300))):
        outs = D.non_max_suppression(d['out'].clone(), **kw)
        for i, o in enumerate(outs):
            assert torch.equal(o, d[f'nms_{style}_{i}']), (style, i)


def test_fuse_only_top_level_convs():
    """models/yolo.py:315-323 + SURVEY F4: fuse() folds BN only for YAML-level `Conv` (cspcm.Conv)."""
    m = D.Model('yolov5s.yaml')
    n_bn = sum(isinstance(x, torch.nn.BatchNorm2d) for x in m.modules())
    x = torch.rand(1, 3, 64, 64)
    m.eval()
    with torch.no_grad():
        a = m(x)[0]
        m.fuse()
        b = m(x)[0]
    n_bn2 = sum(isinstance(x, torch.nn.BatchNorm2d) for x in m.modules())
    assert 0 < n_bn2 < n_bn
    assert_close(a, b, atol=1e-4, rtol=1e-4)


def test_unknown_module_is_reported():
    cfg = dict(nc=2, depth_multiple=1, width_multiple=1, anchors=3, backbone=[[-1, 1, 'NoSuchBlock', [64]]],
               head=[[[0], 1, 'Detect', ['nc', 'anchors']]])
    with pytest.raises(NotImplementedError):
        D.Model(cfg)


def test_aliases_and_pickle_roundtrip(tmp_path):
    import sys
    saved = {k: v for k, v in sys.modules.items() if k == 'models' or k.startswith('models.') or k == 'utils' or k.startswith('utils.')}
    for k in saved:
        del sys.modules[k]
    try:
        D.install_aliases()
        import models.common as mc
        import models.yolo as my
        assert my.Model is D.Model and mc.CA is mc.CoorAttention
        m = D.Model('yolov5s.yaml').eval()
        m(torch.rand(1, 3, 64, 64))
        torch.save({'model': m}, tmp_path / 'ck.pt')
        m2 = torch.load(tmp_path / 'ck.pt', weights_only=False)['model']
        assert state_digest(m2.state_dict()) == state_digest(m.state_dict())
    finally:
        for k in list(sys.modules):
            if k == 'models' or k.startswith('models.') or k == 'utils' or k.startswith('utils.'):
                del sys.modules[k]
        sys.modules.update(saved)


def test_cuda_path_fails_loudly_without_library(monkeypatch):
    from dma_yolo_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'SO', _lib.SO.with_name('missing.so'))
    with pytest.raises(D.DmayError):
        _lib.lib()


def test_tdetect_config_builds_and_runs_on_cpu():
    """CASPD_ODRTA.yaml (SPD-Conv backbone, C3CA neck, four-level anchor-free TDetect head): parse_model,
    the stride probe and bias_init of the TDetect branch (models/yolo.py:173-180, 438-439), eval output layout."""
    from dma_yolo_b200.models.detect_t import TDetect
    m = Y.Model('CASPD_ODRTA.yaml', nc=6)
    head = m.model[-1]
    assert isinstance(head, TDetect) and m.stride.tolist() == [4., 8., 16., 32.]
    assert float(head.cv2[0][-1].bias[0]) == 1.0
    m.eval()
    with torch.no_grad():
        y, (feats, box, cls) = m(torch.rand(1, 3, 64, 64))
    a = 16 * 16 + 8 * 8 + 4 * 4 + 2 * 2
    assert y.shape == (1, 4 + 6, a) and box.shape == (1, 64, a) and cls.shape == (1, 6, a) and len(feats) == 4
    assert torch.isfinite(y).all() and float(y[:, 4:].min()) >= 0.0 and float(y[:, 4:].max()) <= 1.0


_YAML_GOLD = __import__('json').load(open(__import__('pathlib').Path(__file__).parent / 'golden' / 'yaml_digests.json'))


@pytest.mark.parametrize('name', sorted(_YAML_GOLD))
def test_every_reference_yaml_builds_like_the_reference(name):
    """Drop-in claim, literally: every models/*.yaml and models/hub/*.yaml the UNMODIFIED reference can build (65 of
    its 70 files; tests/golden/yaml_digests.json is written by oracle/make_golden_yaml_digests.py from the executed
    reference) builds here with the same seeded state_dict, key for key and bit for bit; the files the reference
    itself cannot build (missing constructor arguments, inconsistent channel counts, hub/anchors.yaml) fail here too.
    Building runs the stride probe, i.e. a CPU train-mode forward of every module (models/yolo.py:161-170)."""
    g = _YAML_GOLD[name]
    path = Y.CFG_DIR / (name + '.yaml')
    if not g['ok']:
        if path.exists():
            with pytest.raises(Exception):
                torch.manual_seed(0)
                Y.Model(str(path))
        return
    torch.manual_seed(0)
    m = Y.Model(str(path))
    assert sum(p.numel() for p in m.parameters()) == g['params']
    assert len(m.state_dict()) == g['keys']
    assert state_digest(m.state_dict()) == g['digest']


# ---- host logic of the lazy concat / space_to_depth folds (weights only; the GEMMs are -m gpu) ---------------------------
def _unpacked_weight(pk):
    """ConvPack.w [Cout_pad][kh][kw][Cin_pad] bf16 -> torch conv weight [Cout, Cin, kh, kw] fp32."""
    return pk.w.float()[:pk.cout, :, :, :pk.cin].permute(0, 3, 1, 2).contiguous()


def test_space_to_depth_fold_weights_are_the_stride2_kernel():
    """get_conv_pack(spd=True): the 1x1 weight over the 4C channels of space_to_depth(x) re-laid as a 2x2 / stride-2 kernel
    over x gives the same convolution (models/common.py:1457-1458 channel order: block q = dy + 2 dx)."""
    import torch.nn.functional as F
    from dma_yolo_b200.models import common as C
    torch.manual_seed(0)
    c, co = 16, 32
    conv = C.Conv(4 * c, co, 1, 1).eval()
    conv.conv.weight.data = conv.conv.weight.data.bfloat16().float()
    x = torch.randn(2, c, 12, 10)
    spd = torch.cat([x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]], 1)
    pk = C.get_conv_pack(conv, 'conv@spd', conv.conv, conv.bn, torch.device('cpu'), spd=True)
    assert (pk.kh, pk.kw, pk.stride, pk.pad, pk.cin) == (2, 2, 2, 0, c)
    ref = F.conv2d(spd, conv.conv.weight)
    got = F.conv2d(x, _unpacked_weight(pk), stride=2)
    assert torch.allclose(got, ref, atol=1e-5, rtol=1e-5)
    plain = C.get_conv_pack(conv, 'conv', conv.conv, conv.bn, torch.device('cpu'))
    assert torch.equal(pk.scale, plain.scale) and torch.equal(pk.bias, plain.bias)     # the folded BN is untouched


def test_lazy_concat_weight_folding_and_column_split():
    """get_conv_pack(colscale=..., cols=..., plain=...): BiFPN weights folded into the input-channel ranges of a 1x1 weight
    (fp32, before the bf16 rounding) and the column split of VCat.split -- W.cat(w_i x_i) == sum_i (w_i W_i).x_i."""
    import torch.nn.functional as F
    from dma_yolo_b200.models import common as C
    torch.manual_seed(1)
    cs, co = (64, 128, 64), 48
    conv = C.Conv(sum(cs), co, 1, 1).eval()
    ws = (0.7, 1.3, 0.4)
    colscale = tuple(zip(cs, ws))
    xs = [torch.randn(1, c, 6, 5) for c in cs]
    ref = F.conv2d(torch.cat([w * x for w, x in zip(ws, xs)], 1), conv.conv.weight)
    pk = C.get_conv_pack(conv, 'conv@vcat', conv.conv, conv.bn, torch.device('cpu'), colscale)
    got = F.conv2d(torch.cat(xs, 1), _unpacked_weight(pk))
    assert torch.allclose(got, ref, atol=2e-2, rtol=2e-2)                   # bf16 rounding of the folded weights
    # split: first part at its own (low) resolution without BN, the rest with it
    lo = C.get_conv_pack(conv, 'conv@lo', conv.conv, conv.bn, torch.device('cpu'), colscale, ((0, 64),), True)
    hi = C.get_conv_pack(conv, 'conv@hi', conv.conv, conv.bn, torch.device('cpu'), colscale, ((64, 192), (192, 256)), False)
    assert lo.cin == 64 and hi.cin == 192
    assert torch.all(lo.scale[:co] == 1) and torch.all(lo.bias == 0) and torch.equal(hi.scale, pk.scale) and torch.equal(hi.bias, pk.bias)
    both = F.conv2d(xs[0], _unpacked_weight(lo)) + F.conv2d(torch.cat(xs[1:], 1), _unpacked_weight(hi))
    assert torch.allclose(both, got, atol=1e-5, rtol=1e-5)
    # the cache is keyed by the fold: another weight tuple rebuilds the pack
    pk2 = C.get_conv_pack(conv, 'conv@vcat', conv.conv, conv.bn, torch.device('cpu'), tuple(zip(cs, (1.0, 1.0, 2.0))))
    assert not torch.equal(pk2.w, pk.w)
