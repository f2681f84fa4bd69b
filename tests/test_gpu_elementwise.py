"""-m gpu: memory-bound kernels vs the oracle, through the C-ABI.  Pure data movement is bit-exact."""
import pytest
import torch
import torch.nn.functional as F

from oracle import blocks as O
from tests.util import assert_close, load_golden

pytestmark = pytest.mark.gpu


def bf(t):
    return t.bfloat16().float()


def cuda_act(t):
    from dma_yolo_b200 import ops
    return ops.as_act(t.cuda())


def back(t):
    return t.float().cpu().contiguous()


def test_layout_roundtrip_and_launch_counter():
    import dma_yolo_b200 as D
    from dma_yolo_b200 import ops
    n0 = D.launch_count()
    x = bf(torch.randn(2, 24, 5, 7))
    a = ops.as_act(x.cuda())
    assert ops.is_nhwc(a) and a.dtype == torch.bfloat16
    assert torch.equal(back(a), x)
    assert torch.equal(ops.to_nchw(a).cpu(), x)
    y = bf(torch.randn(3, 19, 4, 6))                      # channels not a multiple of 8 -> padded slab view
    b = ops.as_act(y.cuda())
    assert b.shape == y.shape and torch.equal(back(b), y)
    assert D.launch_count() > n0


@pytest.mark.parametrize('shape', [(2, 16, 8, 6), (1, 64, 20, 20), (3, 8, 2, 2)])
def test_spd_bit_exact(shape):
    from dma_yolo_b200 import ops
    x = bf(torch.randn(*shape))
    y = ops.spd(cuda_act(x))
    assert torch.equal(back(y), O.space_to_depth(x))


def test_spd_golden_and_into_slab():
    from dma_yolo_b200 import ops
    d, _, ins = load_golden('spd')
    x = ins[0]
    n, c, h, w = x.shape
    slab = ops.empty_nhwc(n, 4 * c + 16, h // 2, w // 2, 'cuda')
    slab.zero_()
    ops.spd(cuda_act(x), out=slab[:, 8:8 + 4 * c])
    assert torch.equal(back(slab[:, 8:8 + 4 * c]), d['out'])
    assert float(slab[:, :8].abs().sum()) == 0 and float(slab[:, 8 + 4 * c:].abs().sum()) == 0


def test_adconcat_golden():
    from dma_yolo_b200.models import common as C
    for name, cls in (('adconcat2', C.AdConcat2), ('adconcat3', C.AdConcat3)):
        d, sd, ins = load_golden(name)
        m = cls()
        m.load_state_dict(sd)
        m = m.cuda().eval()
        y = m([t.cuda() for t in ins])
        assert_close(back(y), d['out'], atol=1e-2, rtol=1e-2, what=name)


def test_adconcat_fused_upsample_and_concat_exact():
    from dma_yolo_b200 import ops
    a = bf(torch.randn(2, 16, 5, 7))
    b = bf(torch.randn(2, 32, 10, 14))
    c = bf(torch.randn(2, 8, 10, 14))
    ref = torch.cat([F.interpolate(a, scale_factor=2, mode='nearest'), b, c], 1)
    y = ops.concat([ops.Up(cuda_act(a), 1), cuda_act(b), cuda_act(c)])
    assert torch.equal(back(y), ref)
    w = (0.25, 0.5, 2.0)                                   # powers of two: products exact in bf16
    y = ops.adconcat([ops.Up(cuda_act(a), 1), cuda_act(b), cuda_act(c)], w)
    ref = torch.cat([w[0] * F.interpolate(a, scale_factor=2, mode='nearest'), w[1] * b, w[2] * c], 1)
    assert torch.equal(back(y), ref)


def test_upsample_golden_bit_exact():
    from dma_yolo_b200 import ops
    d, _, ins = load_golden('upsample2')
    assert torch.equal(back(ops.upsample(cuda_act(ins[0]), 2)), d['out'])


def test_adapt_add_golden():
    from dma_yolo_b200.models import common as C
    d, sd, ins = load_golden('adapt_add2')
    m = C.Adapt_Add2()
    m.load_state_dict(sd)
    y = m.cuda().eval()([t.cuda() for t in ins])
    assert_close(back(y), d['out'], atol=1e-2, rtol=1e-2, what='adapt_add2')


@pytest.mark.parametrize('shape,r', [((2, 16, 38, 38), 4), ((1, 32, 16, 12), 4), ((1, 8, 19, 23), 4), ((1, 8, 9, 9), 2)])
def test_avgpool(shape, r):
    from dma_yolo_b200 import ops
    x = bf(torch.randn(*shape))
    y = ops.avgpool(cuda_act(x), r)
    assert_close(back(y), F.avg_pool2d(x, r, r), atol=1e-2, rtol=1e-2, what='avgpool')


@pytest.mark.parametrize('H,W', [(38, 38), (16, 12), (19, 23), (40, 30), (8, 8)])
def test_scconv_gate_vs_oracle(H, W):
    from dma_yolo_b200 import ops
    g = torch.Generator().manual_seed(H * 100 + W)
    x = bf(torch.randn(2, 16, H, W, generator=g))
    k3 = bf(torch.randn(2, 16, H, W, generator=g))
    k2 = bf(torch.randn(2, 16, H // 4, W // 4, generator=g))
    y = ops.scconv_gate(cuda_act(x), cuda_act(k3), cuda_act(k2))
    ref = O.scconv_gate(x, k3, k2)
    assert_close(back(y), ref, atol=1e-2, rtol=1e-2, what='gate')
    # the nearest-neighbour index map must agree with ATen's F.interpolate exactly
    ref2 = k3 * torch.sigmoid(x + F.interpolate(k2, (H, W)))
    assert_close(ref, ref2, atol=1e-6, rtol=1e-6, what='oracle index map')


def test_input_prep_spd_and_pad():
    from dma_yolo_b200 import ops
    x = torch.rand(2, 3, 12, 8)
    y = ops.input_prep(x.cuda(), spd=True, cpad=16)
    ref = O.space_to_depth(bf(x))
    assert torch.equal(back(y)[:, :12], ref) and float(y[:, 12:].abs().sum()) == 0
    u8 = (x * 255).round().to(torch.uint8)
    y = ops.input_prep(u8.cuda(), spd=False, cpad=16, mul=1 / 255)
    assert_close(back(y)[:, :3], u8.float() / 255, atol=4e-3, rtol=0, what='u8 prep')
