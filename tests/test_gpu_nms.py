"""-m gpu: batched NMS through the C-ABI.  Keep indices / output rows are BIT-EXACT vs the reference
(`tests/golden/nms_*.npz` = outputs of the executed reference `non_max_suppression`) and vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import nms as ON
from tests.test_oracle_golden import STYLES
from tests.util import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('fixture', ['nms_random', 'nms_clustered'])
@pytest.mark.parametrize('style', list(STYLES))
def test_nms_golden_bit_exact(fixture, style):
    import dma_yolo_b200 as D
    d, _, _ = load_golden(fixture)
    kw = dict(STYLES[style])
    if fixture == 'nms_clustered' and style == 'classes':
        kw['classes'] = [0, 2]
    outs = D.non_max_suppression(d['pred'].cuda(), **kw)           # whole batch in one call
    assert len(outs) == d['pred'].shape[0]
    for i, o in enumerate(outs):
        ref = d[f'{style}_{i}']
        assert tuple(o.shape) == tuple(ref.shape), (style, i, tuple(o.shape), tuple(ref.shape))
        assert torch.equal(o.cpu(), ref), (style, i)


def test_nms_truncation_over_max_nms():
    import dma_yolo_b200 as D
    d, _, _ = load_golden('nms_truncate')
    o = D.non_max_suppression(d['pred'].cuda(), 0.001, 0.6, multi_label=True, max_det=300)[0]
    assert torch.equal(o.cpu(), d['val_0'])


def test_nms_known_answers_through_api():
    """tie stability, iou == thr kept, double-vs-float threshold, degenerate boxes (SURVEY 8c KATs)."""
    import dma_yolo_b200 as D

    def run(boxes_xyxy, scores, thr):
        b = torch.tensor(boxes_xyxy, dtype=torch.float32)
        xywh = torch.stack([(b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2, b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]], 1)
        pred = torch.zeros(1, len(b), 6)
        pred[0, :, :4] = xywh
        pred[0, :, 4] = 1.0
        pred[0, :, 5] = torch.tensor(scores)
        out = D.non_max_suppression(pred.cuda(), 0.01, thr)[0].cpu()
        ref = ON.non_max_suppression(pred.numpy(), 0.01, thr)[0]
        assert np.array_equal(out.numpy(), ref)
        return out

    o = run([[0, 0, 10, 10], [100, 100, 110, 110], [0, 0, 10, 10], [200, 200, 210, 210]], [0.5] * 4, 0.5)
    assert o.shape[0] == 3 and o[:, 0].tolist() == [0., 100., 200.]
    assert run([[0, 0, 2, 2], [0, 0, 2, 1]], [0.9, 0.8], 0.5).shape[0] == 2
    assert run([[0, 0, 2, 1], [1, 0, 3, 1]], [0.9, 0.8], 1 / 3).shape[0] == 1
    assert run([[5, 5, 5, 9], [5, 5, 5, 9]], [0.9, 0.8], 0.5).shape[0] == 2


@pytest.mark.parametrize('style', ['detect', 'val'])
def test_nms_class_offset_rounding_and_big_random(style):
    """80 classes: offsets up to 79*4096 round the coordinates in fp32 before IoU; 25,200 rows."""
    import dma_yolo_b200 as D
    g = torch.Generator().manual_seed(21)
    pred = torch.rand(2, 25200, 85, generator=g)
    pred[..., :2] *= 640
    pred[..., 2:4] = pred[..., 2:4] * 60 + 4
    pred[..., 4] = pred[..., 4] ** 4              # fewer candidates (keeps the numpy oracle to seconds)
    pred[..., 5:] = pred[..., 5:] ** 3
    kw = dict(STYLES[style])
    outs = D.non_max_suppression(pred.cuda(), **kw)
    for i in range(2):
        ref = ON.non_max_suppression(pred[i:i + 1].numpy(), **kw)[0]
        assert np.array_equal(outs[i].cpu().numpy(), ref), (style, i)


def cfg5_pred(n=256, seed=1):
    """BASELINE cfg-5 / SURVEY 8d: rand(256, 25200, 15) with pixel-scale xywh; every row passes conf 0.001."""
    pred = torch.rand(n, 25200, 15, generator=torch.Generator().manual_seed(seed))
    pred[..., :2] *= 640
    pred[..., 2:4] = pred[..., 2:4] * 60 + 4
    return pred


@pytest.mark.parametrize('style', ['detect', 'val'])
def test_nms_cfg5_stress_full_size(style):
    """cfg-5 at its full size (256 x 25,200 x 15, ~252 k multi-label rows per image truncated to the top 30,000 in the
    val style): the whole batch in one call; 8 of the 256 images compared BIT-EXACT with the oracle
    (utils/general.py:633-725), every image checked for the size-independent properties (descending scores, count <=
    max_det, finite boxes)."""
    import dma_yolo_b200 as D
    pred = cfg5_pred()
    kw = dict(STYLES[style])
    outs = D.non_max_suppression(pred.cuda(), **kw)
    assert len(outs) == 256
    for o in outs:
        o = o.cpu()
        assert o.shape[0] <= kw.get('max_det', 300) and o.shape[0] > 0
        assert torch.isfinite(o).all()
        assert bool((o[1:, 4] <= o[:-1, 4]).all())
    for i in (0, 1, 37, 100, 128, 200, 254, 255):
        ref = ON.non_max_suppression(pred[i:i + 1].numpy(), **kw)[0]
        assert np.array_equal(outs[i].cpu().numpy(), ref), (style, i)


def test_nms_empty_and_ragged():
    import dma_yolo_b200 as D
    pred = torch.zeros(3, 100, 9)
    pred[1, :5, 4] = 0.9
    pred[1, :5, 5] = 0.8
    pred[1, :5, :4] = torch.tensor([[10, 10, 4, 4], [11, 11, 4, 4], [50, 50, 6, 6], [90, 10, 3, 3], [10, 90, 3, 3]]).float()
    outs = D.non_max_suppression(pred.cuda(), 0.25, 0.45)
    assert [tuple(o.shape) for o in outs] == [(0, 6), tuple(ON.non_max_suppression(pred[1:2].numpy(), 0.25, 0.45)[0].shape), (0, 6)]
    assert np.array_equal(outs[1].cpu().numpy(), ON.non_max_suppression(pred[1:2].numpy(), 0.25, 0.45)[0])
    outs = D.non_max_suppression(torch.zeros(2, 10, 7).cuda(), 0.25, 0.45)
    assert all(tuple(o.shape) == (0, 6) for o in outs)


def test_nms_labels_autolabelling_path():
    import dma_yolo_b200 as D
    d, _, _ = load_golden('nms_random')
    pred = d['pred'][:2, :500].contiguous()
    labels = [torch.tensor([[1., 100., 100., 30., 40.], [3., 300., 200., 50., 20.]]), torch.zeros(0, 5)]
    outs = D.non_max_suppression(pred.cuda(), 0.25, 0.45, labels=[l.cuda() for l in labels])
    ref = ON.non_max_suppression(pred.numpy(), 0.25, 0.45, labels=[l.numpy() for l in labels])
    for o, r in zip(outs, ref):
        assert np.array_equal(o.cpu().numpy(), r)


def test_fused_decode_filter_equals_dense_path():
    """LazyPred (decode fused with the confidence filter) and the materialised dense tensor must give the
    same detections: identical decode arithmetic in both kernels."""
    import dma_yolo_b200 as D
    from dma_yolo_b200.models import yolo as Y
    torch.manual_seed(0)
    det = Y.Detect(nc=10, anchors=[[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]],
                   ch=(32, 64, 128))
    det.stride = torch.tensor([8., 16., 32.])
    det.anchors /= det.stride.view(-1, 1, 1)
    for mi in det.m:
        mi.weight.data.mul_(3.0)
        mi.bias.data.normal_(-1.0, 1.0)
    det = det.cuda().eval()
    det.stride = det.stride.cuda()
    xs = [torch.randn(3, 32, 20, 12).cuda(), torch.randn(3, 64, 10, 6).cuda(), torch.randn(3, 128, 5, 3).cuda()]
    with torch.no_grad():
        kws = (dict(conf_thres=0.25, iou_thres=0.45, max_det=1000), dict(conf_thres=0.001, iou_thres=0.6, multi_label=True))
        pred, _ = det(xs)
        fused_all = [D.non_max_suppression(pred, **kw) for kw in kws]
        assert pred._dense is None, 'fused path must not materialise the dense prediction'
        for kw, fused in zip(kws, fused_all):
            dense = D.non_max_suppression(pred.dense().clone(), **kw)
            oracle = ON.non_max_suppression(pred.dense().cpu().numpy(), **kw)
            for f, dn, o in zip(fused, dense, oracle):
                assert torch.equal(f, dn)
                assert np.array_equal(dn.cpu().numpy(), o)
            assert sum(len(f) for f in fused) > 0


def _quantised_levels(N, nc, shapes, seed, step):
    """Detect logits on a coarse grid of values: confidences collide massively, so equal sort keys are the rule."""
    from dma_yolo_b200 import ops
    g = torch.Generator().manual_seed(seed)
    na, no = 3, 5 + nc
    ld = ops.round_up(na * no, 8)
    levels = []
    for (ny, nx), stride in zip(shapes, (8., 16., 32.)):
        lg = (torch.randn(N, ny, nx, ld, generator=g) * 2.0 / step).round() * step
        anchors = [(stride * (1 + a), stride * (2 + a)) for a in range(na)]
        levels.append(ops.DetectLevel(logits=lg.cuda().contiguous(), stride=stride, anchors_px=anchors, ny=ny, nx=nx, ld=ld))
    return levels, na, no


@pytest.mark.parametrize('kw', [dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300),
                                dict(conf_thres=0.25, iou_thres=0.45, max_det=1000),
                                dict(conf_thres=0.05, iou_thres=0.5, multi_label=True, classes=[1, 4, 6], max_det=100),
                                dict(conf_thres=0.05, iou_thres=0.5, agnostic=True, max_det=50)])
def test_fused_filter_ties_match_ordered_path(kw):
    """Single-pass fused filter (tile look-back compaction) must reproduce the count/scan/write path bit for bit,
    also when thousands of candidates share a score — the stable sort then exposes any ordering slip
    (ragged tiles: 13x7, 5x3 pixels)."""
    from dma_yolo_b200 import ops
    levels, na, no = _quantised_levels(3, 7, [(20, 12), (13, 7), (5, 3)], seed=5, step=0.5)
    fo, fc = ops.nms_batched(None, kw['conf_thres'], kw['iou_thres'], levels=levels, na=na, nc=no - 5,
                             **{k: v for k, v in kw.items() if k not in ('conf_thres', 'iou_thres')})
    dense = ops.detect_decode(levels, na, no)
    do, dc = ops.nms_batched(dense, kw['conf_thres'], kw['iou_thres'],
                             **{k: v for k, v in kw.items() if k not in ('conf_thres', 'iou_thres')})
    assert torch.equal(fc, dc)
    assert int(fc.sum()) > 0
    assert torch.equal(fo, do)
    ref = ON.non_max_suppression(dense.cpu().numpy(), **kw)
    for i, r in enumerate(ref):
        assert np.array_equal(fo[i, :int(fc[i])].cpu().numpy(), r)


def test_fused_filter_capacity_overflow_reruns():
    from dma_yolo_b200 import ops
    levels, na, no = _quantised_levels(2, 7, [(20, 12), (13, 7), (5, 3)], seed=9, step=0.25)
    a = ops.nms_batched(None, 0.001, 0.6, levels=levels, na=na, nc=no - 5, multi_label=True)
    for k in list(ops._FUSED_CAP):
        ops._FUSED_CAP[k] = 64          # far too small: the kernel counts everything, the wrapper repeats
    b = ops.nms_batched(None, 0.001, 0.6, levels=levels, na=na, nc=no - 5, multi_label=True)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


@pytest.mark.parametrize('levels', [3, 40], ids=['three-score-values', 'forty-score-values'])
def test_topk_preselection_is_exact_with_ties(levels):
    """> max_nms candidates per image with massively tied scores: the radix-select + ordered compaction in front of the
    sort must give the same detections as sorting everything (and as the oracle's stable truncation)."""
    from dma_yolo_b200 import ops
    g = torch.Generator().manual_seed(17 + levels)
    pred = torch.rand(3, 9000, 11, generator=g)
    pred[..., :2] *= 640
    pred[..., 2:4] = pred[..., 2:4] * 50 + 4
    pred[..., 4] = 1.0
    pred[..., 5:] = (torch.randint(1, levels + 1, (3, 9000, 6), generator=g).float() / (levels + 1))
    pred[2, 4000:] = 0                                     # third image: fewer than max_nms candidates
    kw = dict(multi_label=True, max_det=300)
    assert ops.TOPK_SELECT
    a, ac = ops.nms_batched(pred.cuda(), 0.001, 0.6, **kw)
    ops.TOPK_SELECT = False
    try:
        b, bc = ops.nms_batched(pred.cuda(), 0.001, 0.6, **kw)
    finally:
        ops.TOPK_SELECT = True
    assert torch.equal(ac, bc) and torch.equal(a, b)
    ref = ON.non_max_suppression(pred[:1].numpy(), 0.001, 0.6, multi_label=True, max_det=300)[0]
    assert np.array_equal(a[0, :int(ac[0])].cpu().numpy(), ref)


def _continuous_levels(N, nc, shapes, seed, scale, special=False):
    from dma_yolo_b200 import ops
    g = torch.Generator().manual_seed(seed)
    na, no = 3, 5 + nc
    ld = ops.round_up(na * no, 8)
    levels = []
    for (ny, nx), stride in zip(shapes, (8., 16., 32.)):
        lg = torch.randn(N, ny, nx, ld, generator=g) * scale
        if special:   # saturating, infinite and NaN logits, objectness right at the threshold
            flat = lg.view(-1)
            idx = torch.randint(0, flat.numel(), (flat.numel() // 50,), generator=g)
            vals = torch.tensor([40.0, -40.0, float('inf'), float('-inf'), float('nan'), 17.5, 9.99, 10.01])
            flat[idx] = vals[torch.randint(0, len(vals), (len(idx),), generator=g)]
        anchors = [(stride * (1 + a), stride * (2 + a)) for a in range(na)]
        levels.append(ops.DetectLevel(logits=lg.cuda().contiguous(), stride=stride, anchors_px=anchors, ny=ny, nx=nx, ld=ld))
    return levels, na, no


@pytest.mark.parametrize('nc', [2, 10, 80, 96, 100])
@pytest.mark.parametrize('thr', [0.0, 0.001, 0.25, 0.9])
@pytest.mark.parametrize('special', [False, True], ids=['finite', 'inf-nan-saturated'])
def test_fused_filter_candidates_match_dense_filter(nc, thr, special):
    """Candidate lists of the single-pass fused filter against the three-launch filter on the materialised dense
    prediction, element for element and in order (keys, boxes, confidences, classes) — multi-label and best-class.
    nc <= 96 runs the thread-per-row kernel whose logit pre-filter must never drop a class the exact test would keep
    (thresholds from 0 to 0.9, saturating / infinite / NaN logits); nc = 100 runs the warp-per-row kernel."""
    from dma_yolo_b200 import ops
    levels, na, no = _continuous_levels(2, nc, [(19, 11), (10, 6), (5, 3)], seed=nc + int(thr * 1000), scale=3.0, special=special)
    dense = ops.detect_decode(levels, na, no)
    for multi in (True, False):
        a = ops.filter_candidates(None, thr, multi_label=multi, levels=levels, na=na, nc=nc)
        b = ops.filter_candidates(dense, thr, multi_label=multi)
        assert torch.equal(a['img_counts'], b['img_counts']), (nc, thr, multi)
        n = int(b['img_counts'].sum())
        assert torch.equal(a['keys'][:n], b['keys'][:n])
        ca, cb = a['cand'][:n], b['cand'][:n]
        assert torch.equal(ca.view(torch.int32), cb.view(torch.int32))


def _pad_levels(levels, na, no):
    """The same logits with every anchor's row padded to a multiple of 4 floats (the layout the Detect head GEMM writes)."""
    from dma_yolo_b200 import ops
    pitch = ops.round_up(no, 4)
    out = []
    for lv in levels:
        n = lv.logits.shape[0]
        ld = ops.round_up(na * pitch, 4)
        lg = torch.full((n, lv.ny, lv.nx, ld), 123.0, device=lv.logits.device)       # pad words must never be read as data
        lg[..., :na * pitch].view(n, lv.ny, lv.nx, na, pitch)[..., :no] = lv.logits[..., :na * no].reshape(n, lv.ny, lv.nx, na, no)
        out.append(ops.DetectLevel(logits=lg, stride=lv.stride, anchors_px=lv.anchors_px, ny=lv.ny, nx=lv.nx, ld=ld, pitch=pitch))
    return out


@pytest.mark.parametrize('nc', [2, 4, 10, 80, 96])
@pytest.mark.parametrize('thr', [0.0, 0.001, 0.25])
@pytest.mark.parametrize('special', [False, True], ids=['finite', 'inf-nan-saturated'])
def test_padded_row_filter_matches_unpadded(nc, thr, special):
    """16-byte-aligned anchor rows (row pitch 4n floats, 16-byte staging, LDS.128 scan) against the unpadded layout:
    candidate keys / boxes / confidences / classes identical bit for bit and in order; the dense decode agrees too."""
    from dma_yolo_b200 import ops
    levels, na, no = _continuous_levels(2, nc, [(19, 11), (10, 6), (5, 3)], seed=7 * nc + int(thr * 1000), scale=3.0, special=special)
    padded = _pad_levels(levels, na, no)
    da, db = ops.detect_decode(levels, na, no), ops.detect_decode(padded, na, no)
    assert torch.equal(da.view(torch.int32), db.view(torch.int32))
    for multi in (True, False):
        a = ops.filter_candidates(None, thr, multi_label=multi, levels=levels, na=na, nc=nc)
        b = ops.filter_candidates(None, thr, multi_label=multi, levels=padded, na=na, nc=nc)
        assert torch.equal(a['img_counts'], b['img_counts']), (nc, thr, multi)
        n = int(a['img_counts'].sum())
        assert torch.equal(a['keys'][:n], b['keys'][:n])
        assert torch.equal(a['cand'][:n].view(torch.int32), b['cand'][:n].view(torch.int32))


@pytest.mark.parametrize('nc', [2, 10, 15, 80])
@pytest.mark.parametrize('thr', [0.0, 0.001, 0.25])
def test_dense_rows_filter_matches_three_launch_filter(nc, thr):
    """Dense predictions through the thread-per-row single-pass filter (KIND 2) against the count / scan / write filter:
    identical candidates in identical order, multi-label and best-class, including NaN / inf rows and rows whose tile is
    not 16-byte aligned (R * no odd)."""
    from dma_yolo_b200 import ops
    g = torch.Generator().manual_seed(31 * nc + int(thr * 1000))
    for R in (777, 1280):
        pred = torch.rand(3, R, 5 + nc, generator=g)
        pred[..., :2] *= 640
        pred[..., 2:4] = pred[..., 2:4] * 60 + 4
        pred[..., 4] = pred[..., 4] ** 2
        flat = pred.view(-1)
        idx = torch.randint(0, flat.numel(), (flat.numel() // 200,), generator=g)
        vals = torch.tensor([float('nan'), float('inf'), -1.0, 0.0, 1.0, 1e-3])
        flat[idx] = vals[torch.randint(0, len(vals), (len(idx),), generator=g)]
        pred = pred.cuda()
        for multi in (True, False):
            a = ops.filter_candidates(pred, thr, multi_label=multi, rows_kernel=True)
            b = ops.filter_candidates(pred, thr, multi_label=multi, rows_kernel=False)
            assert torch.equal(a['img_counts'], b['img_counts']), (nc, thr, multi, R)
            n = int(b['img_counts'].sum())
            assert torch.equal(a['keys'][:n], b['keys'][:n])
            assert torch.equal(a['cand'][:n].view(torch.int32), b['cand'][:n].view(torch.int32))


def test_detect_head_padded_rows_equal_unpadded():
    """Detect.forward_b200 with the head rows padded to 4n floats (default) against the unpadded layout: same raw
    outputs, same dense prediction, same detections."""
    import dma_yolo_b200 as D
    from dma_yolo_b200.models import yolo as Y
    torch.manual_seed(0)
    det = Y.Detect(nc=80, anchors=[[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]],
                   ch=(64, 128, 256))
    det.stride = torch.tensor([8., 16., 32.])
    det.anchors /= det.stride.view(-1, 1, 1)
    for mi in det.m:
        mi.weight.data.mul_(3.0)
        mi.bias.data.normal_(-1.0, 1.0)
    det = det.cuda().eval()
    det.stride = det.stride.cuda()
    xs = [torch.randn(3, 64, 20, 12).cuda(), torch.randn(3, 128, 10, 6).cuda(), torch.randn(3, 256, 5, 3).cuda()]
    kw = dict(conf_thres=0.001, iou_thres=0.6, multi_label=True)
    res = {}
    for pad in (True, False):
        Y.PAD_HEAD_ROWS = pad
        try:
            with torch.no_grad():
                pred, raw = det(list(xs))
                assert (pred._levels[0].pitch == 88) == pad
                dets = D.non_max_suppression(pred, **kw)
                res[pad] = (pred.dense().clone(), [r.clone() for r in raw], [d.clone() for d in dets])
        finally:
            Y.PAD_HEAD_ROWS = True
    assert torch.equal(res[True][0], res[False][0])
    for a, b in zip(res[True][1], res[False][1]):
        assert a.shape == b.shape and torch.equal(a, b)
    for a, b in zip(res[True][2], res[False][2]):
        assert torch.equal(a, b)
    assert sum(len(d) for d in res[True][2]) > 0


@pytest.mark.parametrize('single_cls', [False, True])
def test_val_match_device_equals_host_process_batch(single_cls):
    """`process_batch` + `scale_coords` of a whole batch on the device (dmay_val_match) against the host body that is
    pinned to the executed reference (tests/golden/metrics.npz): same `correct` matrix, image by image — letterboxed
    geometry, images without labels / without detections, labels of classes nobody predicts."""
    from dma_yolo_b200 import val as PV
    from dma_yolo_b200.utils.general import scale_coords, xywh2xyxy
    g = torch.Generator().manual_seed(17)
    B, max_det, nc, H, W = 5, 40, 4, 320, 416
    iouv = torch.linspace(0.5, 0.95, 10)
    shapes = [((240, 400), ((0.8, 0.8), (8.0, 64.0))), ((320, 416), ((1.0, 1.0), (0.0, 0.0))), ((640, 832), ((0.5, 0.5), (0.0, 0.0))),
              ((300, 300), ((1.0666667, 1.0666667), (48.0, 0.0))), ((320, 416), ((1.0, 1.0), (0.0, 0.0)))]
    targets, padded, counts = [], torch.zeros(B, max_det, 6), torch.zeros(B, dtype=torch.int32)
    for b in range(B):
        m = 0 if b == 1 else int(torch.randint(3, 12, (1,), generator=g))
        cxy = torch.rand(m, 2, generator=g) * torch.tensor([W - 80., H - 80.]) + 40
        wh = torch.rand(m, 2, generator=g) * 60 + 12
        cls = torch.randint(0, nc, (m, 1), generator=g).float()
        targets.append(torch.cat((torch.full((m, 1), float(b)), cls, cxy, wh), 1))
        n = 0 if b == 3 else int(torch.randint(10, max_det, (1,), generator=g))
        rows = []
        for k in range(n):
            if m and torch.rand(1, generator=g) < 0.7:
                j = int(torch.randint(0, m, (1,), generator=g))
                c = cxy[j] + torch.randn(2, generator=g) * 2.5
                s = wh[j] * (0.85 + 0.3 * torch.rand(2, generator=g))
                cl = cls[j, 0] if torch.rand(1, generator=g) < 0.8 else torch.randint(0, nc, (1,), generator=g).float()[0]
            else:
                c = torch.rand(2, generator=g) * torch.tensor([W * 1., H * 1.])
                s = torch.rand(2, generator=g) * 50 + 6
                cl = torch.randint(0, nc, (1,), generator=g).float()[0]
            rows.append(torch.stack([c[0] - s[0] / 2, c[1] - s[1] / 2, c[0] + s[0] / 2, c[1] + s[1] / 2, torch.rand(1, generator=g)[0], cl]))
        if n:
            d = torch.stack(rows)
            padded[b, :n] = d[d[:, 4].argsort(descending=True)]
        counts[b] = n
    targets = torch.cat(targets, 0)          # (image, cls, xywh in network-input pixels)
    cor_d, tcls = PV.match_batch_device(padded.cuda(), counts.cuda(), targets, shapes, (H, W), iouv, single_cls)
    cor_d = cor_d.cpu().bool()
    for b in range(B):
        n = int(counts[b])
        pred = padded[b, :n].clone()
        if single_cls:
            pred[:, 5] = 0
        labels = targets[targets[:, 0] == b, 1:]
        assert tcls[b] == labels[:, 0].tolist()
        if n == 0 or len(labels) == 0:
            assert not cor_d[b].any()
            continue
        predn = pred.clone()
        scale_coords((H, W), predn[:, :4], shapes[b][0], shapes[b][1])
        tbox = xywh2xyxy(labels[:, 1:5])
        scale_coords((H, W), tbox, shapes[b][0], shapes[b][1])
        ref = PV.process_batch(predn, torch.cat((labels[:, 0:1], tbox), 1), iouv)
        assert torch.equal(cor_d[b, :n], ref), b
        assert not cor_d[b, n:].any()
    assert cor_d.any()


@pytest.mark.parametrize('maker,step', [('continuous', 0.0), ('quantised', 0.5)], ids=['continuous', 'massive-ties'])
def test_fused_prethreshold_for_candidate_rich_logits_is_exact(maker, step):
    """Detect logits (padded rows, multi-label) with many times max_nms candidates per image: after the first call has seen
    that, the histogram pre-selection (dmay_nms_fused_prethreshold) makes the filter write only what the top-max_nms
    selection can keep.  Detections must be IDENTICAL to the call without it and to the oracle -- also with massively tied
    scores (every candidate of the threshold bin is kept) and with a class filter."""
    from dma_yolo_b200 import ops
    nc = 10
    shapes = [(40, 24), (20, 12), (10, 6)]
    if maker == 'continuous':
        levels, na, no = _continuous_levels(3, nc, shapes, seed=31, scale=2.0)
    else:
        levels, na, no = _quantised_levels(3, nc, shapes, seed=32, step=step)
    levels[0].logits[2] -= 9.0          # third image: (almost) no candidates on its largest level
    padded = _pad_levels(levels, na, no)
    dense = ops.detect_decode(levels, na, no)
    for extra in (dict(), dict(classes=[1, 4, 6])):
        kw = dict(multi_label=True, max_det=300, max_nms=400, **extra)
        ops._FUSED_PRETHR.clear()
        ops._FUSED_CAP.clear()
        a, ac = ops.nms_batched(None, 0.001, 0.6, levels=padded, na=na, nc=nc, **kw)       # plain: sees > 4 x max_nms candidates
        assert any(ops._FUSED_PRETHR.values()), 'the candidate-rich case was not detected'
        n0 = ops.launch_count() if hasattr(ops, 'launch_count') else None
        b, bc = ops.nms_batched(None, 0.001, 0.6, levels=padded, na=na, nc=nc, **kw)       # with the pre-selection
        c, cc = ops.nms_batched(None, 0.001, 0.6, levels=padded, na=na, nc=nc, **kw)       # ... and with its own capacity guess
        assert torch.equal(ac, bc) and torch.equal(a, b) and torch.equal(ac, cc) and torch.equal(a, c)
        ref = ON.non_max_suppression(dense.cpu().numpy(), 0.001, 0.6, multi_label=True, max_det=300, max_nms=400,
                                     **({'classes': extra['classes']} if extra else {}))
        for i, r in enumerate(ref):
            assert np.array_equal(b[i, :int(bc[i])].cpu().numpy(), r), i
    ops._FUSED_PRETHR.clear()
    ops._FUSED_CAP.clear()
