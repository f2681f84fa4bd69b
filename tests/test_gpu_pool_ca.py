"""-m gpu: SPPF pool cascade (bit-exact) and CoordAtt vs oracle/golden."""
import pytest
import torch
import torch.nn.functional as F

from oracle import blocks as O
from tests.util import assert_close, load_golden

pytestmark = pytest.mark.gpu


def bf(t):
    return t.bfloat16().float()


def back(t):
    return t.float().cpu().contiguous()


@pytest.mark.parametrize('shape', [(2, 64, 20, 20), (1, 16, 48, 48), (3, 8, 5, 3), (1, 128, 12, 9), (1, 8, 120, 120)])
def test_pool_cascade_bit_exact(shape):
    from dma_yolo_b200 import ops
    x = bf(torch.randn(*shape))
    n, c, h, w = shape
    slab = ops.empty_nhwc(n, 4 * c, h, w, 'cuda')
    slab[:, :c].copy_(ops.as_act(x.cuda()))
    ops.sppf_pool3(slab[:, :c], slab[:, c:2 * c], slab[:, 2 * c:3 * c], slab[:, 3 * c:], 5)
    y1, y2, y3 = O.maxpool_cascade(x, 5)
    got = back(slab)
    assert torch.equal(got[:, :c], x)
    assert torch.equal(got[:, c:2 * c], y1) and torch.equal(got[:, 2 * c:3 * c], y2) and torch.equal(got[:, 3 * c:], y3)
    for k in (5, 9, 13):
        assert torch.equal(back(ops.maxpool_s1(ops.as_act(x.cuda()), k)), F.max_pool2d(x, k, 1, k // 2))


def test_pool_golden_48():
    from dma_yolo_b200 import ops
    d, _, ins = load_golden('maxpool_cascade_48')
    x = ops.as_act(ins[0].cuda())
    ys = [ops.empty_nhwc(*x.shape, 'cuda') for _ in range(3)]
    ops.sppf_pool3(x, *ys, 5)
    for y, k in zip(ys, ('y1', 'y2', 'y3')):
        assert torch.equal(back(y), d[k])


@pytest.mark.parametrize('name', ['coordatt_7x5', 'coordatt_20x20'])
def test_coordatt_golden(name):
    from dma_yolo_b200.models import common as C
    d, sd, ins = load_golden(name)
    c = ins[0].shape[1]
    m = C.CoorAttention(c, c)
    m.load_state_dict(sd)
    m.bn1.eps = 1e-3
    y = m.cuda().eval()(ins[0].cuda())
    assert_close(back(y), d['out'], atol=1e-2, rtol=1e-2, what=name)


@pytest.mark.parametrize('shape', [(2, 1024, 20, 20), (1, 64, 48, 40), (1, 128, 160, 96), (2, 72, 44, 300), (1, 64, 320, 320)])
def test_coordatt_stages_vs_oracle(shape):
    """pooled means and gates (fp32 side outputs) against the restatement, H != W and large planes (W >= 256: the one-pass
    pool with per-band column partials; 72 channels: a partial last channel group)."""
    from dma_yolo_b200 import ops
    from dma_yolo_b200.models import common as C
    n, c, h, w = shape
    torch.manual_seed(c + h)
    m = C.CoorAttention(c, c).eval()
    m.bn1.running_mean.normal_(0, 0.2)
    m.bn1.running_var.uniform_(0.5, 1.5)
    x = bf(torch.randn(*shape))
    pk = ops.pack_coordatt(m.conv1, m.bn1, m.conv_h, m.conv_w, 'cuda')
    y, pooled, gates = ops.coordatt(ops.as_act(x.cuda()), pk, return_gates=True)
    ref_pool = torch.cat([x.mean(3), x.mean(2)], 2).permute(0, 2, 1)          # [n, h+w, c]
    assert_close(pooled.cpu(), ref_pool, atol=1e-5, rtol=1e-5, what='pooled')
    with torch.no_grad():
        ref = m(x)
    assert_close(back(y), ref, atol=1e-2, rtol=1e-2, what='coordatt out')


def test_coordatt_fast_path_is_repeatable_without_side_outputs():
    """Two-launch path (pool + partial hidden layer with a last-CTA reduction, then gates + apply; taken when the channel
    count is not a multiple of 128): the arrival tickets reset themselves, so repeated calls on one workspace agree bit
    for bit and match the 3-stage result."""
    from dma_yolo_b200 import ops
    from dma_yolo_b200.models import common as C
    torch.manual_seed(3)
    m = C.CoorAttention(192, 192).eval()
    x = ops.as_act(bf(torch.randn(3, 192, 20, 12)).cuda())
    pk = ops.pack_coordatt(m.conv1, m.bn1, m.conv_h, m.conv_w, 'cuda')
    y1 = ops.coordatt(x, pk).clone()
    y2 = ops.coordatt(x, pk).clone()
    y3, _, _ = ops.coordatt(x, pk, return_gates=True)
    assert torch.equal(y1, y2) and torch.equal(y1, y3)
    with torch.no_grad():
        ref = m(back(x))
    assert_close(back(y1), ref, atol=1e-2, rtol=1e-2, what='coordatt fast path')


def _ref_gates(m, x):
    """fp32 gates of the module, [n, h+w, c] (rows 0..h-1: a_h, rows h..: a_w) -- models/common.py:1183-1207."""
    import torch.nn.functional as F
    n, c, h, w = x.shape
    with torch.no_grad():
        y = torch.cat([x.mean(3, keepdim=True), x.mean(2, keepdim=True).permute(0, 1, 3, 2)], 2)
        y = F.hardswish(m.bn1(m.conv1(y)))
        yh, yw = torch.split(y, [h, w], 2)
        a_h = torch.sigmoid(m.conv_h(yh))[..., 0]                        # [n, c, h]
        a_w = torch.sigmoid(m.conv_w(yw.permute(0, 1, 3, 2)))[:, :, 0]   # [n, c, w]
    return torch.cat([a_h, a_w], 2).permute(0, 2, 1).contiguous()


@pytest.mark.parametrize('shape', [(3, 1024, 20, 20), (2, 256, 20, 12), (2, 128, 7, 5), (5, 512, 24, 24), (70, 1024, 20, 15),
                                   (1, 1024, 1, 1), (2, 384, 9, 30), (2, 72, 13, 11), (1, 64, 48, 40)])
def test_coordatt_mma_fast_path(shape):
    """Two-launch path with the hidden layer and the gates on mma.sync (fp32 operands split into bf16 hi + lo, three
    products): bit-repeatable, gates within 3e-5 of the fp32 module (the split keeps ~16 mantissa bits), output within
    tolerance.  Covers Cm = 8/16/32, H != W, a partial last channel group (C = 72), position counts that are not a
    multiple of the 16-row MMA tile and tiles that straddle the h / w gate boundary."""
    from dma_yolo_b200 import ops
    from dma_yolo_b200.models import common as C
    n, c, h, w = shape
    torch.manual_seed(c + h + w)
    m = C.CoorAttention(c, c).eval()
    m.bn1.running_mean.normal_(0, 0.2)
    m.bn1.running_var.uniform_(0.5, 1.5)
    xc = bf(torch.randn(*shape))
    x = ops.as_act(xc.cuda())
    pk = ops.pack_coordatt(m.conv1, m.bn1, m.conv_h, m.conv_w, 'cuda')
    y1 = ops.coordatt(x, pk).clone()
    y2 = ops.coordatt(x, pk).clone()
    y3, pooled, gates = ops.coordatt(x, pk, return_gates=True)
    assert torch.equal(y1, y2) and torch.equal(y1, y3)
    assert_close(gates.cpu(), _ref_gates(m, xc), atol=3e-5, rtol=0, what='gates')
    with torch.no_grad():
        ref = m(xc)
    assert_close(back(y1), ref, atol=1e-2, rtol=1e-2, what='coordatt fast path')


def test_coordatt_strided_slab():
    """Input and output are channel slices of wider NHWC slabs (ldx, ldy > C), as inside a C3 concat slab."""
    from dma_yolo_b200 import ops
    from dma_yolo_b200.models import common as C
    torch.manual_seed(11)
    m = C.CoorAttention(256, 256).eval()
    slab_in = ops.as_act(bf(torch.randn(2, 640, 12, 16)).cuda())
    slab_out = ops.as_act(torch.zeros(2, 512, 12, 16).cuda())
    x = slab_in[:, 128:384]
    pk = ops.pack_coordatt(m.conv1, m.bn1, m.conv_h, m.conv_w, 'cuda')
    ops.coordatt(x, pk, out=slab_out[:, 256:])
    with torch.no_grad():
        ref = m(back(x))
    assert_close(back(slab_out[:, 256:]), ref, atol=1e-2, rtol=1e-2, what='coordatt slab')
    assert float(slab_out[:, :256].abs().max()) == 0.0
