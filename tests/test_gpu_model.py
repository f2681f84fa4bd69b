"""-m gpu: whole detection forward on the kernel path (Model -> Detect -> non_max_suppression)."""
import numpy as np
import pytest
import torch
import yaml

from oracle import blocks as O
from tests.util import assert_close, load_golden

pytestmark = pytest.mark.gpu


def layer_outputs(m, x):
    """per-layer outputs of the kernel path (fp32 NCHW on CPU)."""
    m._trace = []
    with torch.no_grad():
        pred, raw = m(x)
    outs, m._trace = m._trace[:-1], None
    from dma_yolo_b200.ops import Up
    outs = [(o.materialize() if isinstance(o, Up) else o).float().cpu() for o in outs]
    return pred, raw, outs


@pytest.mark.parametrize('cfg', ['yolov5s', 'ablation-ca-scconv-sppfcspc-bifpn'])
def test_model_layerwise_vs_oracle(cfg):
    """Each layer of the CUDA path, fed by the CUDA path's own previous layers, vs the fp32 oracle fed by the
    oracle's previous layers.  bf16 rounding compounds with depth on a random-init net, so the
    check is on the normalised error per layer (||err|| / ||ref||), with the measured numbers reported."""
    import dma_yolo_b200 as D
    from dma_yolo_b200.models import yolo as Y
    from dma_yolo_b200.utils.calib import build_calibrated, state_digest
    d, _, ins = load_golden('model_' + cfg)
    m = build_calibrated(cfg + '.yaml', seed=0)
    assert state_digest(m.state_dict()) == str(d['digest'])
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    strides = m.stride.tolist()
    x = ins[0]
    cfgd = yaml.safe_load(open(Y.CFG_DIR / (cfg + '.yaml')))
    with torch.no_grad():
        ref_pred, _, ref_layers = O.forward_model(cfgd, sd, x, strides)
    m = m.cuda().eval()
    n0 = D.launch_count()
    pred, raw, outs = layer_outputs(m, x.cuda())
    assert D.launch_count() - n0 > 20
    rel = []
    for i, (o, r) in enumerate(zip(outs, ref_layers)):
        assert o.shape == r.shape, (i, o.shape, r.shape)
        rel.append(float((o - r).norm() / (r.norm() + 1e-12)))
    print('per-layer relative L2 error:', ' '.join(f'{v:.4f}' for v in rel))
    assert max(rel[:3]) < 2e-2, rel          # the first layers see almost no compounding
    assert max(rel) < 0.25, rel
    dense = pred.cpu()
    assert dense.shape == d['out'].shape
    assert torch.isfinite(dense).all()


def test_model_end_to_end_nms_runs_and_is_consistent():
    """forward -> fused decode+filter -> NMS on the kernel path; the detections must equal the oracle NMS applied to
    the kernel path's own decoded predictions (bit-exact keep given identical candidate boxes)."""
    import dma_yolo_b200 as D
    from dma_yolo_b200.utils.calib import build_calibrated
    from oracle import nms as ON
    m = build_calibrated('yolov5s.yaml', seed=0).cuda().eval()
    x = torch.rand(4, 3, 96, 128, generator=torch.Generator().manual_seed(5))   # non-square
    with torch.no_grad():
        pred, _ = m(x.cuda())
        for kw in (dict(conf_thres=0.25, iou_thres=0.45, max_det=1000), dict(conf_thres=0.001, iou_thres=0.6, multi_label=True)):
            dets = D.non_max_suppression(pred, **kw)
            ref = ON.non_max_suppression(pred.dense().cpu().numpy(), **kw)
            for a, b in zip(dets, ref):
                assert np.array_equal(a.cpu().numpy(), b)


def test_half_and_uint8_inputs_and_fuse():
    """Input dtypes and model dtypes the callers use (val.py --half, detect.py) all reach the same kernels.
    bf16-exact inputs make fp32 / fp16 / uint8 feeds bit-identical; .half() and .fuse() re-round the weights,
    which a random-init net amplifies, so those are checked on the normalised error."""
    import dma_yolo_b200 as D
    from dma_yolo_b200.utils.calib import build_calibrated
    m = build_calibrated('yolov5s.yaml', seed=0).cuda().eval()
    u8 = torch.randint(0, 256, (2, 3, 64, 64), dtype=torch.uint8)
    x = (u8.float() / 255).bfloat16().float().cuda()
    rel = lambda a, b: float((a - b).norm() / b.norm())
    with torch.no_grad():
        a = m(x)[0].dense().clone()
        b = m(x.half())[0].dense().clone()          # bf16-exact values survive fp16
        assert torch.equal(a, b)
        u = m(u8.cuda())[0].dense().clone()         # uint8 is normalised by the prep kernel (x/255)
        assert rel(u, a) < 0.05
        m.half()
        c = m(x.half())[0].dense().clone()
        m.float().fuse()
        e = m(x)[0].dense().clone()
    assert rel(c, a) < 0.15, rel(c, a)
    assert rel(e, a) < 0.15, rel(e, a)
