"""The oracle (CPU restatement) is pinned against outputs of the executed reference (tests/golden)."""
import numpy as np
import pytest
import torch

from oracle import blocks as O
from oracle import nms as ON
from tests.util import assert_close, load_golden

TOL = dict(atol=2e-5, rtol=2e-5)


def run_block(name):
    d, sd, ins = load_golden(name)
    x = ins[0]
    if name.startswith('conv_k3s1'): y = O.conv_bn_act(x, sd, '', 3, 1)
    elif name.startswith('conv_k3s2'): y = O.conv_bn_act(x, sd, '', 3, 2)
    elif name.startswith('conv_k1'): y = O.conv_bn_act(x, sd, '', 1, 1)
    elif name.startswith('conv_stem'): y = O.conv_bn_act(x, sd, '', 6, 2, 2)
    elif name.startswith('conv_c64'): y = O.conv_bn_act(x, sd, '', 3, 1)
    elif name == 'bottleneck': y = O.bottleneck(x, sd, '')
    elif name == 'c3_n2': y = O.c3(x, sd, '', 2, True)
    elif name == 'c3_n1_noshortcut': y = O.c3(x, sd, '', 1, False)
    elif name.startswith('coordatt'): y = O.coordatt(x, sd, '')
    elif name == 'spd': y = O.space_to_depth(x)
    elif name in ('scconv_38', 'scconv_16x12'): y = O.scconv(x, sd, '', 2)
    elif name == 'scconv_19x23_s1': y = O.scconv(x, sd, '', 1)
    elif name in ('adconcat2', 'adconcat3'): y = O.adconcat(ins, sd['w'])
    elif name == 'adapt_add2': y = O.adapt_add2(ins, sd['w'])
    elif name == 'adapt_add3': y = O.adapt_add3(ins, sd, '')
    elif name == 'sppf_20': y = O.sppf(x, sd, '')
    elif name == 'sppfcspc_12x9': y = O.sppfcspc(x, sd, '')
    elif name == 'spp': y = O.spp(x, sd, '')
    elif name == 'sppcspc': y = O.sppcspc(x, sd, '')
    elif name == 'swin_layer_16x24': y = O.swin_layer(x, sd, '', 2, 8, 0)
    elif name == 'swin_layer_shift_16x24': y = O.swin_layer(x, sd, '', 2, 8, 4)
    elif name == 'swin_layer_shift_13x10': y = O.swin_layer(x, sd, '', 2, 8, 4)
    elif name == 'swin_layer_13x10': y = O.swin_layer(x, sd, '', 1, 8, 0)
    elif name == 'c3str_n2_20x12': y = O.c3str(x, sd, '', 2)
    elif name.startswith('horblock'): y = O.horblock(x, sd, '')
    elif name == 'c3hb_n2_16x12': y = O.c3hb(x, sd, '', 2)
    else: raise KeyError(name)
    return y, d['out']


BLOCKS = ['conv_k3s1', 'conv_k3s2_odd', 'conv_k1', 'conv_stem_k6s2p2', 'conv_c64_k3', 'bottleneck', 'c3_n2',
          'c3_n1_noshortcut', 'coordatt_7x5', 'coordatt_20x20', 'spd', 'scconv_38', 'scconv_16x12', 'scconv_19x23_s1',
          'adconcat2', 'adconcat3', 'adapt_add2', 'adapt_add3', 'sppf_20', 'sppfcspc_12x9', 'spp', 'sppcspc',
          'swin_layer_16x24', 'swin_layer_shift_16x24', 'swin_layer_shift_13x10', 'swin_layer_13x10', 'c3str_n2_20x12',
          'horblock_64_12x10', 'horblock_128_9x7', 'c3hb_n2_16x12']


@pytest.mark.parametrize('name', BLOCKS)
def test_block_oracle_matches_reference(name):
    y, ref = run_block(name)
    if name == 'spd':
        assert torch.equal(y, ref)
    else:
        assert_close(y, ref, what=name, **TOL)


def test_maxpool_cascade_identity():
    d, _, ins = load_golden('maxpool_cascade_48')
    y1, y2, y3 = O.maxpool_cascade(ins[0], 5)
    assert torch.equal(y1, d['y1']) and torch.equal(y2, d['y2']) and torch.equal(y3, d['y3'])  # mp5∘mp5 == mp9 ...


def test_detect_oracle():
    d, sd, ins = load_golden('detect_nc4')
    pred, raw = O.detect(ins, sd, '', sd['anchors'], [8., 16., 32.], 4)
    assert_close(pred, d['out'], what='detect pred', atol=1e-4, rtol=1e-5)
    for i, r in enumerate(raw):
        assert_close(r, d[f'raw{i}'], what=f'raw{i}', **TOL)


@pytest.mark.parametrize('name,nc', [('tdetect_nc10_3lv', 10), ('tdetect_nc20_4lv', 20)])
def test_tdetect_oracle_and_host_mirror(name, nc):
    """Anchor-free TDetect head with DFL decoding (models/detect_t.py): the restatement and the product's host mirror
    (torch body on the CPU, same state_dict keys) against the executed reference."""
    from dma_yolo_b200.models.detect_t import TDetect
    d, sd, ins = load_golden(name)
    strides = d['strides'].tolist()
    y, box, cls = O.tdetect(ins, sd, '', nc, strides)
    assert_close(box, d['box'], what='box logits', **TOL)
    assert_close(cls, d['cls'], what='cls logits', **TOL)
    assert_close(y, d['out'], what='decoded', atol=1e-4, rtol=1e-5)
    m = TDetect(nc, tuple(x.shape[1] for x in ins))
    m.stride = d['strides'].clone()
    m.load_state_dict(sd)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.eps = 1e-3
    m.eval()
    with torch.no_grad():
        ym, (_, bm, cm) = m([x.clone() for x in ins])
    assert_close(ym, d['out'], what='host mirror decoded', atol=1e-4, rtol=1e-5)
    assert_close(bm, d['box'], what='host mirror box', **TOL)


STYLES = {'detect': dict(conf_thres=0.25, iou_thres=0.45, max_det=1000),
          'val': dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300),
          'agnostic': dict(conf_thres=0.3, iou_thres=0.5, agnostic=True, max_det=50),
          'classes': dict(conf_thres=0.2, iou_thres=0.45, classes=[1, 3, 7], max_det=300)}


@pytest.mark.parametrize('fixture', ['nms_random', 'nms_clustered'])
@pytest.mark.parametrize('style', list(STYLES))
def test_nms_oracle_bit_exact(fixture, style):
    d, _, _ = load_golden(fixture)
    kw = dict(STYLES[style])
    if fixture == 'nms_clustered' and style == 'classes':
        kw['classes'] = [0, 2]
    pred = d['pred'].numpy()
    for i in range(pred.shape[0]):
        got = ON.non_max_suppression(pred[i:i + 1], **kw)[0]
        ref = d[f'{style}_{i}'].numpy()
        assert got.shape == ref.shape, (style, i, got.shape, ref.shape)
        assert np.array_equal(got, ref), (style, i)


def test_nms_truncation_path():
    d, _, _ = load_golden('nms_truncate')
    got = ON.non_max_suppression(d['pred'].numpy(), 0.001, 0.6, multi_label=True, max_det=300)[0]
    assert np.array_equal(got, d['val_0'].numpy())


def test_nms_known_answers():
    """SURVEY.md 8c KATs (probed on the reference's torchvision CPU op)."""
    import torchvision
    f = np.float32
    cases = [
        (np.array([[0, 0, 10, 10], [100, 100, 110, 110], [0, 0, 10, 10], [200, 200, 210, 210]], f), np.full(4, 0.5, f), 0.5, [0, 1, 3]),
        (np.array([[0, 0, 2, 2], [0, 0, 2, 1]], f), np.array([0.9, 0.8], f), 0.5, [0, 1]),          # iou == thr kept
        (np.array([[0, 0, 2, 1], [1, 0, 3, 1]], f), np.array([0.9, 0.8], f), 1 / 3, [0]),           # double vs float thr
        (np.array([[5, 5, 5, 9], [5, 5, 5, 9]], f), np.array([0.9, 0.8], f), 0.5, [0, 1]),          # 0/0 -> nan -> kept
    ]
    for boxes, scores, thr, want in cases:
        assert ON.nms_reference(boxes, scores, thr).tolist() == want
        tv = torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).tolist()
        assert tv == want


def test_nms_oracle_vs_torchvision_random():
    import torchvision
    g = torch.Generator().manual_seed(3)
    for trial in range(6):
        n = 1500
        xy = torch.rand(n, 2, generator=g) * 300
        wh = torch.rand(n, 2, generator=g) * 80 + 1
        boxes = torch.cat([xy, xy + wh], 1)
        scores = torch.rand(n, generator=g)
        if trial % 2:
            scores = (scores * 20).round() / 20          # many ties
        off = (torch.randint(0, 80, (n, 1), generator=g).float() * 4096) if trial >= 3 else 0
        b = boxes + off                                 # fp32 rounding of class offsets
        for thr in (0.45, 0.6):
            want = torchvision.ops.nms(b, scores, thr).numpy()
            got = ON.nms_reference(b.numpy(), scores.numpy(), thr)
            assert np.array_equal(got, want), (trial, thr)


def test_metrics_oracle_and_host_port_match_the_executed_reference():
    """mAP arithmetic (val.process_batch + utils.metrics.ap_per_class): the numpy oracle AND the product's host-side
    mirror (dma_yolo_b200.val / utils.metrics) against outputs of the executed reference (tests/golden/metrics.npz)."""
    import numpy as np
    import torch
    from dma_yolo_b200 import val as PV
    from dma_yolo_b200.utils import metrics as PM
    from dma_yolo_b200.utils.general import xywh2xyxy
    from oracle import metrics as OM
    from pathlib import Path
    d = np.load(Path(__file__).resolve().parent / 'golden' / 'metrics.npz')
    n_img, S = int(d['n_img']), int(d['S'])
    dets = [d[f'det_{i}'] for i in range(n_img)]
    labs = [d[f'lab_{i}'] for i in range(n_img)]
    iouv = np.linspace(0.5, 0.95, 10).astype(np.float32)
    stats_o, stats_p = [], []
    for i in range(n_img):
        if f'correct_{i}' not in d.files:
            continue
        half = labs[i][:, 3:5] * S / 2
        lab = np.concatenate([labs[i][:, 0:1], labs[i][:, 1:3] * S - half, labs[i][:, 1:3] * S + half], 1).astype(np.float32)
        if len(labs[i]):
            co = OM.process_batch(dets[i], lab, iouv)
            lt = torch.from_numpy(labs[i])
            cp = PV.process_batch(torch.from_numpy(dets[i]), torch.cat((lt[:, 0:1], xywh2xyxy(lt[:, 1:5] * S)), 1),
                                  torch.linspace(0.5, 0.95, 10)).numpy()
            assert np.array_equal(co, d[f'correct_{i}']), i
            assert np.array_equal(cp, d[f'correct_{i}']), i
    mp, mr, m50, m = OM.evaluate(dets, labs, (S, S))
    assert np.allclose([mp, mr, m50, m], d['summary'], atol=1e-9), ([mp, mr, m50, m], d['summary'])
    # product port of ap_per_class on the reference's own `correct` matrices
    tp = np.concatenate([d[f'correct_{i}'] for i in range(n_img) if f'correct_{i}' in d.files], 0)
    conf = np.concatenate([dets[i][:, 4] for i in range(n_img) if f'correct_{i}' in d.files], 0)
    pcls = np.concatenate([dets[i][:, 5] for i in range(n_img) if f'correct_{i}' in d.files], 0)
    tcls = np.concatenate([labs[i][:, 0] for i in range(n_img) if (len(dets[i]) or len(labs[i]))], 0)
    p, r, ap, f1, cls = PM.ap_per_class(tp, conf, pcls, tcls)
    assert np.allclose(ap, d['ap'], atol=1e-9) and np.allclose(p, d['p'], atol=1e-9) and np.allclose(r, d['r'], atol=1e-9)
    assert np.array_equal(cls, d['ap_class'])


def test_letterbox_host_mirror_matches_executed_reference():
    """8f-3: `letterbox` (utils/augmentations.py:91-122) -- resize + pad geometry, ratios and per-side padding, for the
    auto / fixed / scaleFill / no-scale-up branches and non-default stride and colour; pixels bit-exact (same cv2 calls)."""
    from dma_yolo_b200.utils.augmentations import letterbox
    from oracle.make_golden_letterbox import CASES
    z = np.load(__import__('tests.util', fromlist=['GOLD']).GOLD / 'letterbox.npz')
    for i, (h, w, kw) in enumerate(CASES):
        im = z[f'im{i}']
        assert im.shape == (h, w, 3)
        out, ratio, pad = letterbox(im.copy(), **kw)
        assert out.shape == z[f'out{i}'].shape and np.array_equal(out, z[f'out{i}']), (i, kw)
        assert np.allclose([ratio[0], ratio[1], pad[0], pad[1]], z[f'meta{i}'], rtol=0, atol=1e-12), (i, kw)


@pytest.mark.parametrize('tag', ['ablation', 'c3caspd'])
def test_oracle_whole_path_map_equals_reference_val_run(tag):
    """Whole-path pin of the oracle: forward_model (fp32) -> numpy non_max_suppression -> metrics.evaluate on the trained
    checkpoint and the seeded labelled val set reproduces the mAP@0.5:0.95 that the UNMODIFIED reference's val.run
    printed for the same weights and images (tests/golden/conditioned_<tag>.json, oracle/train_conditioned.py) to
    1e-9 — so the 1e-4 mAP criterion of the GPU tests is judged against the reference itself, not against a port."""
    import json
    from pathlib import Path

    from oracle import blocks as O
    from oracle import metrics as OM
    from oracle import synth
    gold = Path(__file__).parent / 'golden'
    ck = torch.load(gold / f'conditioned_{tag}.pt', map_location='cpu')
    info = json.load(open(gold / f'conditioned_{tag}.json'))
    sd = {k: (v.float() if v.is_floating_point() else v) for k, v in ck['state_dict'].items()}
    S = info['size']
    kw = dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)
    dets, labels = [], []
    for b in range(info['val_batches']):
        im, tg = synth.make_batch(info['val_seed'] + b, info['val_b'], S)
        with torch.no_grad():
            pred, _, _ = O.forward_model(ck['cfg'], sd, torch.from_numpy(im).float() / 255, ck['stride'])
        for i in range(im.shape[0]):
            dets.append(ON.non_max_suppression(pred[i:i + 1].numpy(), **kw)[0])
            labels.append(tg[tg[:, 0] == i][:, 1:])
        if b == 0:
            ref0 = np.load(gold / f'conditioned_{tag}_dets.npz')
            for i in range(im.shape[0]):      # the reference's own detections of the first batch, row for row
                assert dets[i].shape == ref0[f'det_{i}'].shape
                assert np.allclose(dets[i], ref0[f'det_{i}'], atol=2e-3, rtol=1e-4), i
    mp, mr, map50, map_ = OM.evaluate(dets, labels, (S, S))
    assert abs(map_ - info['map']) < 1e-9 and abs(map50 - info['map50']) < 1e-9, (map_, info['map'])
    assert abs(mp - info['mp']) < 1e-6 and abs(mr - info['mr']) < 1e-6    # P, R: interpolated at the best-F1 confidence


def test_oracle_tta_matches_reference_forward_augment():
    """oracle.blocks.forward_augment against `model(x, augment=True)` of the executed reference (conditioned checkpoint,
    two 256 x 320 images; tests/golden/tta_conditioned.npz, oracle/make_golden_tta.py): 1e-4 relative."""
    from pathlib import Path

    from oracle import blocks as O
    gold = Path(__file__).parent / 'golden'
    d = np.load(gold / 'tta_conditioned.npz')
    ck = torch.load(gold / 'conditioned_ablation.pt', map_location='cpu')
    sd = {k: (v.float() if v.is_floating_point() else v) for k, v in ck['state_dict'].items()}
    x = torch.from_numpy(d['x'])
    with torch.no_grad():
        out = O.forward_augment(ck['cfg'], sd, x, ck['stride'], nl=len(ck['stride']))
        plain = O.forward_model(ck['cfg'], sd, x, ck['stride'])[0]
    assert tuple(out.shape) == d['out'].shape
    assert np.allclose(plain.numpy(), d['plain'], atol=2e-3, rtol=1e-4)
    assert np.allclose(out.numpy(), d['out'], atol=2e-3, rtol=1e-4)
