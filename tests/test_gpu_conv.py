"""tcgen05 3x3 convolution (csrc/tc_conv.cu) against torch conv2d in float64 on the same inputs, at every shape family
the UNet uses (single / two-image tiles, partial tiles in x and y, concatenated inputs, residual, 1-2 output blocks),
and the whole 2D-network plan against the module it replaces.  Tolerance: 2e-5 of max|reference| (bf16 hi/lo x 3
products drop only the lo*lo term, ~2^-16 per product; the end-to-end logit bar is 1e-4)."""
import warnings

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def ref_conv(x1, x2, w, b, res, relu):
    x = x1 if x2 is None else torch.cat([x1, x2], dim=3)
    y = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1)
    if res is not None:
        y = y + res.double()
    return y.clamp_min(0) if relu else y


@pytest.mark.parametrize('n,h,w,c1,c2,cout,res,relu', [
    (3, 64, 80, 64, 0, 64, True, True),        # layer1
    (2, 128, 160, 64, 64, 64, False, True),    # decoder0: concat, largest tiles count
    (3, 32, 40, 128, 0, 128, True, True),      # layer2
    (3, 16, 20, 256, 0, 256, True, True),      # layer3: partial tile in x
    (5, 8, 10, 512, 0, 512, True, True),       # layer4: two images per tile, odd image count, two output blocks
    (3, 16, 20, 256, 256, 256, False, True),   # decoder3
    (2, 30, 37, 16, 0, 16, False, False),      # ragged everything, no activation: signed outputs
    (1, 5, 3, 32, 16, 48, True, False),
    (7, 8, 8, 16, 0, 32, False, True),
    (1, 16, 8, 16, 0, 64, False, True),        # a single tile: the peer CTA of the pair has none
    (3, 40, 24, 32, 0, 96, True, True),        # 27 tiles: ragged last pair item, 96-column block
    (5, 17, 9, 16, 16, 32, False, False),
])
def test_conv3x3_matches_float64(n, h, w, c1, c2, cout, res, relu):
    from mvpnet_b200 import net2d
    torch.manual_seed(n * 1000 + h)
    dev = 'cuda'
    x1 = torch.randn(n, h, w, c1, device=dev)
    x2 = torch.randn(n, h, w, c2, device=dev) if c2 else None
    wt = torch.randn(cout, c1 + c2, 3, 3, device=dev) / (3.0 * (c1 + c2) ** 0.5)
    b = torch.randn(cout, device=dev)
    r = torch.randn(n, h, w, cout, device=dev) if res else None
    packed, bias = net2d.pack_conv3x3(wt, b)
    P = net2d.Planar.from_nhwc
    want = ref_conv(x1, x2, wt, b, r, relu)
    scale = want.abs().max().item()
    # the split-planar round trip itself: hi + lo carries ~17 mantissa bits
    assert (P(x1).to_nhwc() - x1).abs().max().item() <= 2 ** -16 * x1.abs().max().item()
    got_p = net2d.conv3x3(P(x1), packed, bias, x2=None if x2 is None else P(x2), residual=None if r is None else P(r), relu=relu)
    assert (got_p.n, got_p.h, got_p.w, got_p.c) == (n, h, w, cout)
    err = (got_p.to_nhwc().double() - want).abs().max().item() / scale
    assert err < 2e-5, err
    got_f = net2d.conv3x3(P(x1), packed, bias, x2=None if x2 is None else P(x2), residual=None if r is None else P(r), relu=relu,
                          nhwc_out=True)
    err = (got_f.double() - want).abs().max().item() / scale
    assert err < 2e-5, err
    # the CTA-pair kernel (narrow layers) and the single-CTA kernel issue the same products in the same order: bit-identical
    got_s = net2d.conv3x3(P(x1), packed, bias, x2=None if x2 is None else P(x2), residual=None if r is None else P(r), relu=relu, pair=False)
    assert torch.equal(got_s.data, got_p.data)
    # row-split output (what the fused FeatureAggregation gathers): bf16 (hi, lo) planes, pixel-major rows
    got_r = net2d.conv3x3(P(x1), packed, bias, x2=None if x2 is None else P(x2), residual=None if r is None else P(r), relu=relu,
                          nhwc_out=2)
    assert got_r.dtype == torch.bfloat16 and tuple(got_r.shape) == (2, n, h, w, cout)
    err = (got_r[0].double() + got_r[1].double() - want).abs().max().item() / scale
    assert err < 2e-5, err


def test_conv3x3_errors():
    from mvpnet_b200 import net2d
    x = net2d.Planar.from_nhwc(torch.randn(1, 8, 8, 24, device='cuda'))
    packed, bias = net2d.pack_conv3x3(torch.randn(16, 32, 3, 3), torch.zeros(16))
    with pytest.raises(RuntimeError):
        net2d.conv3x3(x, packed.cuda(), bias.cuda())               # 24 channels: not a multiple of 16 / wrong weight size


def test_fast_unet_matches_module():
    from mvpnet_b200 import net2d, synthetic
    from mvpnet_b200.unet import UNetResNet34
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        net = UNetResNet34(20, p=0.5, pretrained=False)
    synthetic.fill_parameters(net, seed=4)
    net = net.eval().cuda()
    plan = net2d.FastUNetResNet34(net)
    x = torch.randn(5, 3, 120, 160, device='cuda')
    with torch.no_grad():
        want = net.double().features(x.double())
        got = plan.features_nhwc(x).permute(0, 3, 1, 2)
    assert got.shape == want.shape
    err = (got.double() - want).abs().max().item() / want.abs().max().item()
    assert err < 5e-5, err


def _rel(got, want):
    return (got.double() - want).abs().max().item() / want.abs().max().item()


@pytest.mark.parametrize('n,h,w,cin,cout,k', [
    (3, 64, 80, 64, 128, 3), (3, 32, 40, 128, 256, 3), (5, 16, 20, 256, 512, 3),     # the three strided 3x3 convolutions
    (3, 64, 80, 64, 128, 1), (5, 16, 20, 256, 512, 1), (2, 30, 37, 16, 32, 3), (3, 9, 11, 32, 16, 1)])
def test_strided_conv_matches_float64(n, h, w, cin, cout, k):
    from mvpnet_b200 import net2d
    torch.manual_seed(h * 10 + k)
    x = torch.randn(n, h, w, cin, device='cuda')
    wt = torch.randn(cout, cin, k, k, device='cuda') / (k * cin ** 0.5)
    b = torch.randn(cout, device='cuda')
    packed, bias, dy, dx = net2d.pack_conv_taps(wt, b)
    for relu in (True, False):
        got = net2d.conv_general(net2d.Planar.from_nhwc(x), packed, bias, dy, dx, stride=2, relu=relu)
        want = F.conv2d(x.permute(0, 3, 1, 2).double(), wt.double(), b.double(), stride=2, padding=k // 2).permute(0, 2, 3, 1)
        want = want.clamp_min(0) if relu else want
        assert (got.n, got.h, got.w, got.c) == tuple(want.shape)
        assert _rel(got.to_nhwc(), want) < 2e-5


@pytest.mark.parametrize('n,h,w,cin,cout', [(5, 8, 10, 512, 256), (3, 16, 20, 256, 128), (3, 32, 40, 128, 64), (2, 64, 80, 64, 64),
                                            (3, 5, 7, 32, 16)])
def test_deconv2x2_matches_float64(n, h, w, cin, cout):
    from mvpnet_b200 import net2d
    torch.manual_seed(h)
    x = torch.randn(n, h, w, cin, device='cuda')
    wt = torch.randn(cin, cout, 2, 2, device='cuda') / cin ** 0.5
    b = torch.randn(cout, device='cuda')
    packed, bias = net2d.pack_deconv2x2(wt, b)
    got = net2d.deconv2x2(net2d.Planar.from_nhwc(x), packed, bias, relu=True)
    want = F.conv_transpose2d(x.permute(0, 3, 1, 2).double(), wt.double(), b.double(), stride=2).permute(0, 2, 3, 1).clamp_min(0)
    assert (got.n, got.h, got.w, got.c) == tuple(want.shape)
    assert _rel(got.to_nhwc(), want) < 2e-5


def test_stem_and_maxpool_match_float64():
    import mvpnet_b200
    from mvpnet_b200 import net2d
    ext = mvpnet_b200.load_ext()
    torch.manual_seed(3)
    img = torch.randn(3, 3, 48, 64, device='cuda')
    wt = torch.randn(64, 3, 7, 7, device='cuda') / 12.0
    b = torch.randn(64, device='cuda')
    packed, bias, dy, dx = net2d.pack_stem7x7(wt, b)
    x = net2d.Planar(ext.fused_cuda.unfold_stem(img), 3, 48, 64, 32)
    got = net2d.conv_general(x, packed, bias, dy, dx, stride=1, relu=True)
    want = F.conv2d(img.double(), wt.double(), b.double(), padding=3).clamp_min(0)
    assert _rel(got.to_nhwc(), want.permute(0, 2, 3, 1)) < 2e-5
    for h, w in [(48, 64), (16, 20), (15, 9)]:            # 16 -> 8 rows: pair-interleaved output; odd sizes
        t = torch.randn(3, h, w, 16, device='cuda')
        pooled = net2d.maxpool3x3s2(net2d.Planar.from_nhwc(t))
        ref = F.max_pool2d(net2d.Planar.from_nhwc(t).to_nhwc().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
        assert torch.equal(pooled.to_nhwc(), ref.contiguous())
