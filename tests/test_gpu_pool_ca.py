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


@pytest.mark.parametrize('shape', [(2, 1024, 20, 20), (1, 64, 48, 40), (1, 128, 160, 96)])
def test_coordatt_stages_vs_oracle(shape):
    """pooled means and gates (fp32 side outputs) against the restatement, H != W and large planes."""
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
    """Two-launch path (pool + partial hidden layer with a last-CTA reduction, then gates + apply): the arrival
    tickets reset themselves, so repeated calls on one workspace agree bit for bit and match the 3-stage result."""
    from dma_yolo_b200 import ops
    from dma_yolo_b200.models import common as C
    torch.manual_seed(3)
    m = C.CoorAttention(256, 256).eval()
    x = ops.as_act(bf(torch.randn(3, 256, 20, 12)).cuda())
    pk = ops.pack_coordatt(m.conv1, m.bn1, m.conv_h, m.conv_w, 'cuda')
    y1 = ops.coordatt(x, pk).clone()
    y2 = ops.coordatt(x, pk).clone()
    y3, _, _ = ops.coordatt(x, pk, return_gates=True)
    assert torch.equal(y1, y2) and torch.equal(y1, y3)
    with torch.no_grad():
        ref = m(back(x))
    assert_close(back(y1), ref, atol=1e-2, rtol=1e-2, what='coordatt fast path')
