"""-m gpu: the tcgen05 implicit-GEMM conv through the C-ABI vs F.conv2d (fp32, same bf16-rounded operands).

Accumulation is fp32 in TMEM on exact bf16 products, so before the bf16 store the result equals the fp32
reference up to summation order; tolerance = bf16 output rounding (2^-9 relative) + slack: atol=rtol=1e-2."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import assert_close

pytestmark = pytest.mark.gpu


def bf(t):
    return t.bfloat16().float()


def back(t):
    return t.float().cpu().contiguous()


def run_conv(n, cin, h, w, cout, k, s, p, act=1, residual=False, out_fp32=False, bias=False, seed=0, slab_pad=0):
    from dma_yolo_b200 import ops
    g = torch.Generator().manual_seed(seed)
    x = bf(torch.randn(n, cin, h, w, generator=g))
    wt = bf(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5)
    scale = torch.rand(cout, generator=g) + 0.5
    b = torch.randn(cout, generator=g) * 0.3
    bn = torch.nn.BatchNorm2d(cout).eval()
    bn.weight.data, bn.bias.data = scale.clone(), b.clone()
    bn.running_mean.zero_(); bn.running_var.fill_(1.0); bn.eps = 0.0
    conv_bias = bf(torch.randn(cout, generator=g)) if bias else None
    pk = ops.pack_conv(wt, bn=bn, conv_bias=conv_bias, stride=s, pad=p, device='cuda')
    ref = F.conv2d(x, wt, conv_bias, s, p) * scale.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
    if act == 1:
        ref = ref * torch.sigmoid(ref)
    res = None
    if residual:
        res = bf(torch.randn(ref.shape, generator=g))
        ref = ref + res
    xin = x.cuda() if cin % 16 else ops.as_act(x.cuda())
    out = None
    if slab_pad:
        ho, wo = ref.shape[2:]
        slab = ops.empty_nhwc(n, cout + 2 * slab_pad, ho, wo, 'cuda')
        slab.zero_()
        out = slab[:, slab_pad:slab_pad + cout]
    y = ops.conv(xin, pk, act, out=out, residual=None if res is None else ops.as_act(res.cuda()), out_fp32=out_fp32)
    torch.cuda.synchronize()
    if slab_pad:
        assert float(slab[:, :slab_pad].abs().sum()) == 0 and float(slab[:, slab_pad + cout:].abs().sum()) == 0
    return back(y), ref


CASES = [
    # n, cin, h, w, cout, k, s, p
    (1, 64, 8, 8, 64, 1, 1, 0),        # plain GEMM, one tile
    (2, 64, 20, 20, 128, 1, 1, 0),     # M = 800: tail tile
    (1, 128, 16, 16, 256, 1, 1, 0),    # 2 K chunks, N = 256
    (1, 256, 12, 10, 512, 1, 1, 0),    # 2 n-tiles
    (2, 64, 12, 10, 64, 3, 1, 1),      # im2col 3x3
    (1, 64, 20, 20, 64, 3, 2, 1),      # stride 2
    (3, 128, 11, 9, 128, 3, 2, 1),     # odd sizes, stride 2, multi image
    (1, 32, 16, 16, 64, 3, 1, 1),      # Cin = 32 -> 64-byte swizzle
    (1, 16, 16, 16, 32, 3, 1, 1),      # Cin = 16 -> 32-byte swizzle
    (1, 512, 10, 10, 1024, 3, 1, 1),   # K = 4608, 4 n-tiles
    (4, 64, 40, 40, 64, 3, 1, 1),      # 50 m-tiles, persistent loop + double-buffered TMEM
    (1, 64, 5, 5, 48, 1, 1, 0),        # N = 48 (x16 TMEM tail)
    (1, 1024, 20, 20, 512, 1, 1, 0),   # K = 1024
]


@pytest.mark.parametrize('case', CASES, ids=[str(c) for c in CASES])
def test_conv_bn_silu(case):
    y, ref = run_conv(*case)
    assert_close(y, ref, atol=1e-2, rtol=1e-2, what=str(case))


def test_conv_stem_spd_rewrite():
    y, ref = run_conv(2, 3, 64, 48, 64, 6, 2, 2)          # 6x6 s2 p2 on RGB == 3x3 s1 p1 on pixel-unshuffled input
    assert_close(y, ref, atol=1e-2, rtol=1e-2, what='stem')
    y, ref = run_conv(1, 3, 32, 32, 32, 3, 1, 1, seed=3)  # generic small-Cin path (channel padding to 16)
    assert_close(y, ref, atol=1e-2, rtol=1e-2, what='cin3 k3')


def test_conv_residual_and_slab_slice():
    y, ref = run_conv(2, 64, 12, 12, 64, 3, 1, 1, residual=True, slab_pad=16)
    assert_close(y, ref, atol=1e-2, rtol=1e-2, what='residual+slab')


def test_conv_detect_head_fp32_255():
    y, ref = run_conv(2, 128, 10, 10, 255, 1, 1, 0, act=0, out_fp32=True, bias=True)
    assert y.shape[1] == 255
    assert_close(y, ref, atol=2e-4, rtol=2e-4, what='fp32 head')   # fp32 store: only summation order differs


def test_conv_gate_epilogue():
    from dma_yolo_b200 import ops
    from oracle import blocks as O
    g = torch.Generator().manual_seed(5)
    n, c, h, w = 2, 64, 19, 23
    x = bf(torch.randn(n, c, h, w, generator=g))
    wt = bf(torch.randn(c, c, 3, 3, generator=g) / (c * 9) ** 0.5)
    k2 = bf(torch.randn(n, c, h // 4, w // 4, generator=g))
    pk = ops.pack_conv(wt, stride=1, pad=1, device='cuda')
    xa = ops.as_act(x.cuda())
    y = ops.conv(xa, pk, 0, gate=(xa, ops.as_act(k2.cuda())))
    ref = O.scconv_gate(x, F.conv2d(x, wt, None, 1, 1), k2)
    assert_close(back(y), ref, atol=1e-2, rtol=1e-2, what='gate epilogue')


def test_conv_rejects_bad_arguments():
    import dma_yolo_b200 as D
    from dma_yolo_b200 import ops
    pk = ops.pack_conv(torch.randn(64, 64, 1, 1), device='cuda')
    with pytest.raises(D.DmayError):
        ops.conv(torch.randn(1, 32, 4, 4).cuda(), pk)


# ---- kernel path variants that auto-selection only engages at full-size layers: force them at test sizes ----
def run_flags(n, cin, h, w, cout, flags=0, block_n=0, residual=False, gate=False, seed=0):
    from dma_yolo_b200 import ops
    from oracle import blocks as O
    g = torch.Generator().manual_seed(seed)
    x = bf(torch.randn(n, cin, h, w, generator=g))
    wt = bf(torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5)
    pk = ops.pack_conv(wt, stride=1, pad=1, device='cuda')
    ref = F.conv2d(x, wt, None, 1, 1)
    xa = ops.as_act(x.cuda())
    kw = {}
    if gate:
        k2 = bf(torch.randn(n, cout, max(h // 4, 1), max(w // 4, 1), generator=g))
        ref = O.scconv_gate(x, ref, k2)
        kw['gate'] = (xa, ops.as_act(k2.cuda()))
        act = 0
    else:
        ref = ref * torch.sigmoid(ref)
        act = 1
    if residual:
        res = bf(torch.randn(ref.shape, generator=g))
        ref = ref + res
        kw['residual'] = ops.as_act(res.cuda())
    y = ops.conv(xa, pk, act, flags=flags, block_n=block_n, **kw)
    torch.cuda.synchronize()
    return back(y), ref


HALO_SHAPES = [(2, 64, 32, 16, 64), (1, 128, 48, 40, 128), (2, 64, 20, 20, 64), (1, 16, 32, 32, 32), (1, 32, 16, 24, 64),
               (3, 256, 16, 8, 256), (1, 64, 19, 23, 64), (2, 128, 33, 9, 48)]


@pytest.mark.parametrize('shape', HALO_SHAPES, ids=[str(s) for s in HALO_SHAPES])
@pytest.mark.parametrize('flags', [2, 2 | 4], ids=['halo-resident-if-fits', 'halo-streamed-weights'])
def test_conv_halo_paths(shape, flags):
    """3x3 s1 p1 through the halo path (input patch loaded once, taps = shifted UMMA windows), incl. maps the
    16x8 patches do not tile (19x23, 33x9) and every swizzle width (Cin 16 / 32 / 64+)."""
    y, ref = run_flags(*shape, flags=flags)
    assert_close(y, ref, atol=1e-2, rtol=1e-2, what=f'halo {shape} flags={flags}')


@pytest.mark.parametrize('kw', [dict(residual=True), dict(gate=True)], ids=['residual', 'gate'])
@pytest.mark.parametrize('flags', [1, 2], ids=['im2col', 'halo'])
def test_conv_epilogue_variants_on_both_paths(kw, flags):
    y, ref = run_flags(2, 64, 24, 16, 64, flags=flags, seed=9, **kw)
    assert_close(y, ref, atol=1e-2, rtol=1e-2, what=f'{kw} flags={flags}')
    y, ref = run_flags(1, 128, 19, 23, 128, flags=flags, seed=10, **kw)     # M tail + non-tiling map
    assert_close(y, ref, atol=1e-2, rtol=1e-2, what=f'{kw} flags={flags} odd')


def test_conv_cluster_multicast_path():
    """block_n = -2: weight tile fetched in two halves and multicast across a 2-CTA cluster (opt-in experiment)."""
    for shape in [(2, 64, 32, 16, 64), (1, 256, 24, 24, 256), (1, 128, 17, 13, 128)]:
        y, ref = run_flags(*shape, flags=1, block_n=-2)
        assert_close(y, ref, atol=1e-2, rtol=1e-2, what=f'cluster {shape}')


PAIR_SHAPES = [(2, 256, 24, 24, 256), (1, 128, 17, 13, 256), (3, 64, 16, 8, 512), (1, 256, 20, 20, 128), (1, 512, 9, 7, 384)]


@pytest.mark.parametrize('shape', PAIR_SHAPES, ids=[str(s) for s in PAIR_SHAPES])
@pytest.mark.parametrize('kw', [dict(), dict(residual=True), dict(gate=True)], ids=['plain', 'residual', 'gate'])
def test_conv_cta_pair_mode(shape, kw):
    """flags bit8: cta_group::2 — a 2-CTA cluster computes one 256 x block_n tile (each CTA loads its own 128 rows
    of A and half of the weight tile, the even CTA issues the MMAs into both CTAs' TMEM).  Odd tile counts (a
    padding m-tile), M tails, two n-tiles, block_n 128 / 256 and 128-wide last tiles, with every epilogue."""
    if kw.get('gate') and shape[1] != shape[4]:
        pytest.skip('the SCConv gate multiplies by x: needs Cin == Cout')
    y, ref = run_flags(*shape, flags=256 | 1, seed=3, **kw)
    assert_close(y, ref, atol=1e-2, rtol=1e-2, what=f'pair {shape} {kw}')


def test_conv_cta_pair_mode_1x1_and_strided():
    from dma_yolo_b200 import ops
    g = torch.Generator().manual_seed(12)
    for (n, cin, h, w, cout, k, s) in [(2, 1024, 20, 20, 256, 1, 1), (1, 256, 23, 19, 512, 3, 2), (4, 512, 10, 10, 1024, 1, 1)]:
        x = bf(torch.randn(n, cin, h, w, generator=g))
        wt = bf(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5)
        pk = ops.pack_conv(wt, stride=s, pad=k // 2, device='cuda')
        ref = F.conv2d(x, wt, None, s, k // 2)
        ref = ref * torch.sigmoid(ref)
        y = ops.conv(ops.as_act(x.cuda()), pk, 1, flags=256)
        y0 = ops.conv(ops.as_act(x.cuda()), pk, 1, flags=512)
        torch.cuda.synchronize()
        assert_close(back(y), ref, atol=1e-2, rtol=1e-2, what=f'pair {(n, cin, h, w, cout, k, s)}')
        assert_close(back(y), back(y0), atol=1e-2, rtol=1e-2, what='pair vs single-CTA tiles')


def test_conv_weights_resident_plain_path():
    """flags bit11: the whole weight set stays in shared memory on the plain (1x1 / im2col) path and only activation
    tiles are streamed (auto-selected for shallow-K layers with many tiles); checked against the streamed-weights path."""
    from dma_yolo_b200 import ops
    g = torch.Generator().manual_seed(31)
    for (n, cin, h, w, cout, k, s) in [(2, 256, 24, 20, 256, 1, 1), (3, 128, 17, 13, 128, 1, 1), (2, 64, 16, 16, 128, 3, 2),
                                       (1, 64, 9, 7, 48, 1, 1), (2, 32, 12, 12, 64, 3, 1)]:
        x = bf(torch.randn(n, cin, h, w, generator=g))
        wt = bf(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5)
        pk = ops.pack_conv(wt, stride=s, pad=k // 2, device='cuda')
        ref = F.conv2d(x, wt, None, s, k // 2)
        ref = ref * torch.sigmoid(ref)
        y = ops.conv(ops.as_act(x.cuda()), pk, 1, flags=2048 | 1)
        y0 = ops.conv(ops.as_act(x.cuda()), pk, 1, flags=1024 | 1)
        torch.cuda.synchronize()
        assert_close(back(y), ref, atol=1e-2, rtol=1e-2, what=f'resident {(n, cin, h, w, cout, k, s)}')
        assert torch.equal(back(y), back(y0)), 'resident and streamed weights must give identical results'


@pytest.mark.parametrize('shape', [(2, 128, 32, 24, 128), (1, 128, 19, 23, 256), (3, 64, 16, 8, 128), (1, 256, 17, 9, 128)],
                         ids=str)
@pytest.mark.parametrize('kw', [dict(), dict(residual=True), dict(gate=True)], ids=['plain', 'residual', 'gate'])
def test_conv_cta_pair_with_halo(shape, kw):
    """flags bit8 | bit1: CTA-pair tiles on the halo path — each CTA of the pair loads its own 18x10 input patch and half
    of the 9-tap weight set; the even CTA's MMAs read both CTAs' patches through shifted-window descriptors."""
    if kw.get('gate') and shape[1] != shape[4]:
        pytest.skip('the SCConv gate multiplies by x: needs Cin == Cout')
    y, ref = run_flags(*shape, flags=256 | 2, seed=4, **kw)
    assert_close(y, ref, atol=1e-2, rtol=1e-2, what=f'pair+halo {shape} {kw}')


@pytest.mark.parametrize('shape', [(2, 128, 32, 24, 128), (3, 128, 16, 8, 128), (3, 64, 16, 8, 128), (1, 128, 19, 23, 128),
                                   (5, 128, 40, 40, 128), (2, 64, 32, 16, 64), (3, 64, 40, 24, 64)], ids=str)
@pytest.mark.parametrize('kw', [dict(), dict(residual=True), dict(gate=True)], ids=['plain', 'residual', 'gate'])
def test_conv_cta_pair_halo_resident_weights(shape, kw):
    """CTA pair + halo with each CTA's HALF of the 9-tap weight set resident in shared memory (loaded once per CTA, both
    halves credited to the even CTA's barrier) against the same mode with streamed weights (flags bit2): identical
    bits, and both within tolerance of the fp32 reference.  Includes an odd number of m-tiles (zero-filled padding tile
    in the last pair) and a map the 16x8 patches do not tile."""
    if kw.get('gate') and shape[1] != shape[4]:
        pytest.skip('the SCConv gate multiplies by x: needs Cin == Cout')
    y, ref = run_flags(*shape, flags=256 | 2, seed=6, **kw)
    y0, _ = run_flags(*shape, flags=256 | 2 | 4, seed=6, **kw)
    assert torch.equal(y, y0), 'resident and streamed weights must give identical results'
    assert_close(y, ref, atol=1e-2, rtol=1e-2, what=f'pair+halo+resident {shape} {kw}')


@pytest.mark.parametrize('shape', [(2, 128, 10, 10, 255), (1, 256, 13, 7, 45), (3, 64, 20, 12, 30), (1, 512, 5, 3, 255)], ids=str)
def test_conv_fp32_output_staged_equals_direct(shape):
    """Detect-head convs (1x1, bias, fp32 output): the staged 128-byte-swizzled tile + TMA store must write exactly what the
    direct per-lane stores (flags bit14) write, including Cout that is not a multiple of 8 / 32 and an M tail."""
    from dma_yolo_b200 import ops
    n, cin, h, w, cout = shape
    g = torch.Generator().manual_seed(cout + h)
    x = ops.as_act(bf(torch.randn(n, cin, h, w, generator=g)).cuda())
    wt = bf(torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5)
    bias = torch.randn(cout, generator=g)
    pk = ops.pack_conv(wt, conv_bias=bias, stride=1, pad=0, device='cuda')
    y = ops.conv(x, pk, 0, out_fp32=True)
    y0 = ops.conv(x, pk, 0, out_fp32=True, flags=16384)
    torch.cuda.synchronize()
    assert y.dtype == torch.float32 and y.shape[1] == cout and torch.equal(y, y0)
    ref = F.conv2d(back(x), wt, bias)
    assert_close(y.cpu(), ref, atol=2e-4, rtol=2e-4, what=f'fp32 head {shape}')


@pytest.mark.parametrize('cin,cout,hw', [(16, 64, (64, 48)), (64, 64, (40, 24)), (32, 32, (44, 28))])
def test_conv_epilogue_avgpool4_byproduct_is_bit_identical(cin, cout, hw):
    """pool4_out of dmay_conv_bn_act (plain SiLU 3x3 layers with resident weights: the stem in front of an SCConv): the 4x4
    average of the layer's OWN bf16 output, written from the epilogue's staging tile, equals dmay_avgpool of that output bit
    for bit -- also where the 16 x 8 patches overhang the map (44 x 28) -- and the output itself is unchanged."""
    from dma_yolo_b200 import ops
    g = torch.Generator().manual_seed(cin + cout)
    h, w = hw
    x = ops.as_act(torch.randn(48, cin, h, w, generator=g).cuda())      # >= 2 x 148 patches: the halo / resident-weight plan
    pk = ops.pack_conv(torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5, stride=1, pad=1, device='cuda')
    y_plain = ops.conv(x, pk, ops.ACT_SILU)
    y = ops.conv(x, pk, ops.ACT_SILU, pool4=True)
    torch.cuda.synchronize()
    assert torch.equal(y, y_plain)
    pooled = getattr(y, '_dmay_pool4', None)
    assert pooled is not None, 'the layer qualifies for the pooling epilogue (T9 image) but fell back'
    assert torch.equal(pooled, ops.avgpool(y, 4))
    # a layer that does not qualify (residual epilogue) silently runs the plain launch
    r = ops.as_act(torch.randn(48, cout, h, w, generator=g).cuda())
    y2 = ops.conv(x, pk, ops.ACT_SILU, residual=r, pool4=True)
    assert getattr(y2, '_dmay_pool4', None) is None and torch.equal(y2, ops.conv(x, pk, ops.ACT_SILU, residual=r))
