"""-m gpu: drop-in modules on the kernel path vs outputs of the executed reference (tests/golden).

Per-block parity on identical bf16 inputs/weights: single-kernel blocks atol=rtol=1e-2;
blocks chaining several bf16-rounded convs get 3e-2 (each intermediate is re-rounded to bf16: 2^-9
relative per layer on O(1..10) activations)."""
import pytest
import torch

from tests.test_modules_cpu import CTORS, make_module
from tests.util import assert_close, load_golden

pytestmark = pytest.mark.gpu

SINGLE = {'conv_k3s1', 'conv_k3s2_odd', 'conv_k1', 'conv_stem_k6s2p2', 'conv_c64_k3', 'coordatt_7x5', 'coordatt_20x20',
          'adconcat2', 'adconcat3', 'adapt_add2'}


@pytest.mark.parametrize('name', list(CTORS))
def test_block_kernel_path_matches_reference(name):
    import dma_yolo_b200 as D
    m, d, ins = make_module(name)
    m = m.cuda().eval()
    n0 = D.launch_count()
    with torch.no_grad():
        y = m([t.cuda() for t in ins] if len(ins) > 1 else ins[0].cuda())
    torch.cuda.synchronize()
    from dma_yolo_b200 import ops
    if isinstance(y, ops.Up):      # AdConcat / Concat return a lazy concat when a 1x1 consumer could read the parts in place
        y = y.materialize()
        torch.cuda.synchronize()
    assert D.launch_count() > n0, 'no kernel of libdmayolo.so was launched'
    y = y.float().cpu()
    if name == 'spd':
        assert torch.equal(y, d['out'])
        return
    tol = 1e-2 if name in SINGLE else 3e-2
    assert_close(y, d['out'], atol=tol, rtol=tol, what=name)


def test_scconv_unfused_gate_variant():
    from dma_yolo_b200.models import common as C
    m, d, ins = make_module('scconv_16x12')
    m = m.cuda().eval()
    C._State.fuse_scconv_gate = False
    try:
        with torch.no_grad():
            y = m(ins[0].cuda())
    finally:
        C._State.fuse_scconv_gate = True
    assert_close(y.float().cpu(), d['out'], atol=3e-2, rtol=3e-2, what='scconv unfused')


def test_detect_head_lazy_and_dense():
    from dma_yolo_b200.lazy import LazyPred
    from dma_yolo_b200.models import yolo as Y
    d, sd, ins = load_golden('detect_nc4')
    det = Y.Detect(nc=4, anchors=[[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]],
                   ch=(16, 32, 64))
    det.stride = torch.tensor([8., 16., 32.])
    det.load_state_dict(sd, strict=True)
    det = det.cuda().eval()
    det.stride = det.stride.cuda()
    with torch.no_grad():
        pred, raw = det([t.cuda() for t in ins])
    assert isinstance(pred, LazyPred) and tuple(pred.shape) == tuple(d['out'].shape)
    for i, r in enumerate(raw):
        assert_close(r.float().cpu(), d[f'raw{i}'], atol=2e-3, rtol=2e-3, what=f'raw{i}')
    dense = pred.cpu()                                   # any torch op materialises through the decode kernel
    assert_close(dense, d['out'], atol=5e-2, rtol=2e-3, what='decoded pred')
    # decode arithmetic itself: reference formula applied to OUR logits, 1e-4 relative
    z = []
    for i, r in enumerate(raw):
        y = r.float().cpu().sigmoid()
        bs, na, ny, nx, no = y.shape
        yv, xv = torch.meshgrid(torch.arange(ny), torch.arange(nx), indexing='ij')
        grid = torch.stack((xv, yv), 2).view(1, 1, ny, nx, 2).float()
        ag = (sd['anchors'][i] * [8., 16., 32.][i]).view(1, na, 1, 1, 2)
        y[..., 0:2] = (y[..., 0:2] * 2 - 0.5 + grid) * [8., 16., 32.][i]
        y[..., 2:4] = (y[..., 2:4] * 2) ** 2 * ag
        z.append(y.reshape(bs, -1, no))
    assert_close(dense, torch.cat(z, 1), atol=1e-5, rtol=1e-4, what='decode of identical logits')


@pytest.mark.parametrize('shape', [(2, 64, 16, 24, 0), (2, 64, 16, 24, 4), (1, 128, 13, 10, 4), (3, 32, 8, 8, 0), (1, 256, 20, 36, 4)],
                         ids=str)
def test_window_attention_mma_vs_scalar_kernel(shape):
    """The mma.sync window-attention kernel against the scalar fp32 kernel of the same library on identical qkv
    (both against the oracle through the Swin block fixtures): padding, shift, mask and the transposed frame."""
    from dma_yolo_b200 import ops
    from dma_yolo_b200.models import common as C
    n, c, h, w, shift = shape
    torch.manual_seed(c + h + shift)
    layer = C.SwinTransformerLayer(c, num_heads=c // 32, window_size=8, shift_size=shift).eval()
    with torch.no_grad():
        layer.attn.relative_position_bias_table.normal_(0, 0.5)
    qkv = ops.as_act(torch.randn(n, 3 * c, h, w).bfloat16().float().cuda())
    rel = layer.attn.rel_bias().float().cuda().contiguous()
    mask = layer.create_mask(torch.zeros(1), w, h).float().cuda().contiguous() if shift else None
    a = ops.window_attention(qkv, rel, mask, c // 32, shift, layer.attn.scale, variant=0)
    b = ops.window_attention(qkv, rel, mask, c // 32, shift, layer.attn.scale, variant=1)
    torch.cuda.synchronize()
    assert_close(a.float().cpu(), b.float().cpu(), atol=2e-2, rtol=2e-2, what=f'attention mma vs scalar {shape}')


@pytest.mark.parametrize('name,nc', [('tdetect_nc10_3lv', 10), ('tdetect_nc20_4lv', 20)])
def test_tdetect_kernel_path(name, nc):
    """8f-4: TDetect (models/detect_t.py) on the kernel path -- six tcgen05 convs per level (the 1x1 heads with staged fp32
    output), DFL expectation + box arithmetic in fp32 -- against the executed reference's outputs."""
    from dma_yolo_b200.models.detect_t import TDetect
    d, sd, ins = load_golden(name)
    m = TDetect(nc, tuple(x.shape[1] for x in ins))
    m.stride = d['strides'].clone()
    m.load_state_dict(sd)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.eps = 1e-3
    m = m.cuda().eval()
    m.stride = m.stride.cuda()
    with torch.no_grad():
        y, (feats, box, cls) = m([x.cuda() for x in ins])
    assert y.dtype == torch.float32 and y.shape == d['out'].shape
    assert_close(box, d['box'], atol=4e-2, rtol=3e-2, what='box logits')         # three chained bf16 convs
    assert_close(cls, d['cls'], atol=4e-2, rtol=3e-2, what='cls logits')
    assert_close(y[:, 4:], d['out'][:, 4:], atol=1e-2, rtol=1e-2, what='class confidences')
    # boxes: the DFL expectation is a softmax average over 16 bins; tolerance in units of each point's stride
    per_point_stride = torch.cat([torch.full((x.shape[-2] * x.shape[-1],), float(s)) for x, s in zip(ins, d['strides'])])
    err = (y[:, :4].cpu() - d['out'][:, :4]).abs() / per_point_stride.view(1, 1, -1)
    assert float(err.max()) < 0.1, float(err.max())


@pytest.mark.parametrize('name,nc', [('tdetect_nc10_3lv', 10), ('tdetect_nc20_4lv', 20)])
def test_dfl_decode_kernel_equals_reference_arithmetic(name, nc):
    """The one-kernel TDetect tail (DFL softmax expectation + dist2bbox + stride + class sigmoid) against the reference's
    own fp32 arithmetic (DFL module, dist2bbox, make_anchors of models/detect_t.py) applied to the SAME head logits:
    1e-4 relative on the boxes, 1e-6 on the confidences."""
    from dma_yolo_b200.models import detect_t as T
    d, sd, ins = load_golden(name)
    m = T.TDetect(nc, tuple(x.shape[1] for x in ins))
    m.stride = d['strides'].clone()
    m.load_state_dict(sd)
    m = m.cuda().eval()
    m.stride = m.stride.cuda()
    xs = [x.cuda() for x in ins]
    with torch.no_grad():
        y_k, (_, box, cls) = m(list(xs))
        T.DFL_KERNEL = False
        try:
            m.shape = None
            y_t, (_, box2, cls2) = m(list(xs))
        finally:
            T.DFL_KERNEL = True
    assert torch.equal(box, box2) and torch.equal(cls, cls2)
    assert y_k.shape == y_t.shape
    assert_close(y_k[:, :4], y_t[:, :4], atol=1e-3, rtol=1e-4, what='DFL boxes')
    assert_close(y_k[:, 4:], y_t[:, 4:], atol=1e-6, rtol=1e-5, what='class confidences')


# ---- virtual concat: AdConcat2 / AdConcat3 / Concat feeding a 1x1 consumer without writing the concat -------------------
def _vcat_parts(kind, n=2, hw=(24, 40), seed=0):
    from dma_yolo_b200 import ops
    g = torch.Generator().manual_seed(seed)
    h, w = hw
    mk = lambda c, hh, ww: ops.as_act(torch.randn(n, c, hh, ww, generator=g).cuda())
    if kind == 'up2':        # [Up(low-res), same-res]: the head's top-down joins
        return [ops.Up(mk(128, h // 2, w // 2), 1), mk(64, h, w)]
    if kind == 'three':      # AdConcat3: three same-resolution parts of different widths, one a channel slice of a slab
        slab = mk(256, h, w)
        return [mk(64, h, w), slab[:, 64:192], mk(192, h, w)]
    return [mk(128, h, w), mk(128, h, w)]


@pytest.mark.parametrize('kind', ['up2', 'three', 'two'])
def test_conv_over_virtual_concat_is_bit_identical_to_the_written_concat(kind):
    """dmay_conv_bn_act with x / x1 / x2 (weights 1.0: nothing folded) against the same GEMM over the concat written by
    the adconcat kernel: the K loop only changes where a 64-channel chunk is fetched from, so the outputs are EQUAL."""
    from dma_yolo_b200 import ops
    parts = _vcat_parts(kind)
    vc = ops.vcat(parts, (1.0,) * len(parts))
    assert isinstance(vc, ops.VCat) and vc.colscale() is None
    cin = vc.shape[1]
    g = torch.Generator().manual_seed(5)
    for cout in (64, 264, 512):
        pk = ops.pack_conv(torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5, device='cuda')
        y_virtual = ops.conv(vc.sources(), pk, ops.ACT_SILU)
        y_written = ops.conv(vc.materialize(), pk, ops.ACT_SILU)
        torch.cuda.synchronize()
        assert torch.equal(y_virtual, y_written), (kind, cout)
    with pytest.raises(ops.DmayError):       # 3x3 layers cannot walk parts
        ops.conv(vc.sources(), ops.pack_conv(torch.randn(64, cin, 3, 3), pad=1, device='cuda'), ops.ACT_SILU)


@pytest.mark.parametrize('split', [False, True])
@pytest.mark.parametrize('mod', ['AdConcat2', 'AdConcat3', 'Concat'])
def test_c3_reads_a_lazy_concat_in_place(mod, split, monkeypatch):
    """AdConcatN / Concat -> C3 through the layer loop's dispatch: the lazy form (parts read in place, BiFPN weights folded
    into cv1 | cv2's columns in fp32 before the bf16 rounding) against the materialised form (adconcat kernel, then C3).
    Concat folds nothing: equal.  AdConcat rounds w_i * W instead of w_i * x: 1e-2 relative to the output scale."""
    import dma_yolo_b200 as D
    from dma_yolo_b200 import ops
    from dma_yolo_b200.models import common as C
    from dma_yolo_b200.models import yolo as Y
    # split: the `Up` part runs as a low-resolution partial GEMM (fp32 sums) that the main GEMM's epilogue adds before
    # scale / bias / SiLU (ops.VCat.split, opt-in) -- another summation order, so only Concat WITHOUT an Up part stays equal
    monkeypatch.setattr(ops, 'SPLIT_UP', split)
    torch.manual_seed(3)
    parts = _vcat_parts('three' if mod == 'AdConcat3' else 'up2', seed=7)
    cat = getattr(C, mod)(1).cuda().eval()
    if mod != 'Concat':
        cat.w.data = torch.tensor([0.7, 1.3, 0.4][:len(parts)]).cuda()
    cin = sum(p.shape[1] for p in parts)
    c3 = C.C3(cin, 128, n=1, shortcut=False).cuda().eval()
    for m_ in c3.modules():
        if isinstance(m_, torch.nn.BatchNorm2d):
            m_.running_var.data.uniform_(0.5, 1.5)
            m_.running_mean.data.normal_(0, 0.1)
    run = Y.Model._run_layer
    with torch.no_grad():
        lazy = cat(parts)
        assert isinstance(lazy, ops.VCat)
        n0 = D.launch_count()
        y_lazy = run(None, c3, lazy, True)
        n_lazy = D.launch_count() - n0
        assert lazy._mat is None, 'the concat was written although its consumer can read the parts'
        n0 = D.launch_count()
        y_mat = run(None, c3, ops.adconcat(parts, cat._norm_weights() if mod != 'Concat' else (1.0,) * len(parts)), True)
        n_mat = D.launch_count() - n0
    torch.cuda.synchronize()
    a, b = y_lazy.float(), y_mat.float()
    if mod == 'Concat' and not split:
        assert torch.equal(a, b)
    else:
        rel = float((a - b).norm() / b.norm())
        assert rel < 4e-3 and float((a - b).abs().max()) <= 1e-2 * float(b.abs().max()), rel
    assert n_lazy <= n_mat    # the replicate of an `Up` part replaces the adconcat launch; same-resolution parts cost nothing


@pytest.mark.parametrize('cout,hw,up', [(128, (24, 40), 2), (264, (20, 20), 2), (64, (16, 16), 4)])
def test_conv_pre_add_epilogue_matches_torch(cout, hw, up):
    """dmay_conv_bn_act with `pre` (EPI_SILU_PRE): y = silu(scale * (W.x + nearest_up(pre)) + bias) against torch fp32 on
    the same bf16-rounded x / W and the same fp32 partial sums: atol = rtol = 1e-2 (single kernel)."""
    import torch.nn.functional as F
    from dma_yolo_b200 import ops
    g = torch.Generator().manual_seed(11)
    h, w = hw
    cin = 128
    x = torch.randn(2, cin, h, w, generator=g).bfloat16().float()
    wt = (torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5).bfloat16().float()
    bn = torch.nn.BatchNorm2d(cout).eval()
    bn.weight.data.uniform_(0.5, 1.5, generator=g)
    bn.bias.data.normal_(0, 0.2, generator=g)
    bn.running_mean.data.normal_(0, 0.2, generator=g)
    bn.running_var.data.uniform_(0.5, 1.5, generator=g)
    pre = torch.randn(2, cout, h // up, w // up, generator=g)
    with torch.no_grad():
        ref = F.silu(bn(F.conv2d(x, wt) + F.interpolate(pre, scale_factor=up, mode='nearest')))
    pk = ops.pack_conv(wt, bn=bn, device='cuda')
    pre_d = ops.empty_nhwc(2, cout, h // up, w // up, 'cuda', torch.float32, c_alloc=(cout + 7) // 8 * 8)
    pre_d.copy_(pre.cuda())
    y = ops.conv(ops.as_act(x.cuda()), pk, ops.ACT_SILU, pre=pre_d)
    torch.cuda.synchronize()
    assert_close(y.float().cpu(), ref, atol=1e-2, rtol=1e-2, what=f'pre-add epilogue cout={cout}')
    with pytest.raises(ops.DmayError):     # bf16 partial sums are not accepted
        ops.conv(ops.as_act(x.cuda()), pk, ops.ACT_SILU, pre=pre_d.bfloat16())


@pytest.mark.parametrize('c,hw', [(64, (32, 48)), (128, (20, 20)), (16, (8, 8))])
def test_c3_folds_space_to_depth_into_a_stride2_gemm(c, hw):
    """space_to_depth -> C3 through the layer loop's dispatch: cv1 | cv2 (1x1 over 4C channels) run as ONE 2x2 / stride-2 GEMM
    over the un-shuffled tensor (ops.SPDView; the 4C-channel tensor is never written) against the materialised form (dmay_spd,
    then the 1x1 GEMM).  Same products, another summation order over K: 4e-3 in rel-L2, 1e-2 of the output scale."""
    import dma_yolo_b200 as D
    from dma_yolo_b200 import ops
    from dma_yolo_b200.models import common as C
    from dma_yolo_b200.models import yolo as Y
    torch.manual_seed(5)
    h, w = hw
    x = ops.as_act(torch.randn(2, c, h, w).cuda())
    spd = C.space_to_depth().cuda().eval()
    c3 = C.C3(4 * c, 64, n=1).cuda().eval()
    for m_ in c3.modules():
        if isinstance(m_, torch.nn.BatchNorm2d):
            m_.running_var.data.uniform_(0.5, 1.5)
            m_.running_mean.data.normal_(0, 0.1)
    run = Y.Model._run_layer
    with torch.no_grad():
        lazy = run(None, spd, x, True)
        assert isinstance(lazy, ops.SPDView) and tuple(lazy.shape) == (2, 4 * c, h // 2, w // 2)
        y_lazy = run(None, c3, lazy, True)
        assert lazy._mat is None, 'space_to_depth was written although its consumer folds it'
        mat = lazy.materialize()
        ref = torch.cat([x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]], 1)
        assert torch.equal(mat, ref)                      # the lazy object still materialises to the reference's layout
        y_mat = run(None, c3, mat, True)
    torch.cuda.synchronize()
    a, b = y_lazy.float(), y_mat.float()
    rel = float((a - b).norm() / b.norm())
    assert rel < 4e-3 and float((a - b).abs().max()) <= 1e-2 * float(b.abs().max()), rel


def test_conv1x1_reads_a_lazy_concat_and_others_materialise_it():
    """A 1x1 `Conv` after AdConcat reads the parts in place (folded weights); a 3x3 `Conv` gets the written concat (once)."""
    from dma_yolo_b200 import ops
    from dma_yolo_b200.models import common as C
    from dma_yolo_b200.models import yolo as Y
    torch.manual_seed(9)
    parts = _vcat_parts('three', seed=3)
    cat = C.AdConcat3(1).cuda().eval()
    cat.w.data = torch.tensor([1.1, 0.6, 0.9]).cuda()
    cin = sum(p.shape[1] for p in parts)
    run = Y.Model._run_layer
    with torch.no_grad():
        for k in (1, 3):
            conv = C.Conv(cin, 96, k, 1).cuda().eval()
            conv.bn.running_var.data.uniform_(0.5, 1.5)
            lazy = cat(parts)
            y_lazy = run(None, conv, lazy, True)
            assert (lazy._mat is None) == (k == 1)
            y_mat = run(None, conv, ops.adconcat(parts, cat._norm_weights()), True)
            torch.cuda.synchronize()
            a, b = y_lazy.float(), y_mat.float()
            if k == 3:
                assert torch.equal(a, b)
            else:
                assert float((a - b).norm() / b.norm()) < 4e-3
