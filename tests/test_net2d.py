"""Host logic of the tensor-core 2D network plan (mvpnet_b200/net2d.py) on CPU: weight packing round trip and the
wiring of the plan (BatchNorm folding, residuals, concat-free decoder) against UNetResNet34.features, with the CUDA
convolution replaced by a torch emulation that consumes the PACKED weights."""
import warnings

import torch
import torch.nn.functional as F

from mvpnet_b200 import net2d, synthetic


def unpack_taps(packed, cin, g, t, nt):
    x = packed.view(torch.bfloat16).reshape(g // nt, cin // 16, t, 2, 2, nt, 8)          # (nb, c, t, hl, k8, n, e)
    x = x.permute(3, 0, 5, 1, 4, 6, 2).reshape(2, g, cin, t).float()
    return x[0], x[1]


def conv3x3_nt(cout):
    return min(cout, 256)            # mvp_tc_conv3x3_nt


def unpack_conv3x3(packed, cin, cout, nt=None):
    hi, lo = unpack_taps(packed, cin, cout, 9, nt or conv3x3_nt(cout))
    return hi.reshape(cout, cin, 3, 3), lo.reshape(cout, cin, 3, 3)


def emulated_conv(x1, x2, packed, bias, residual, relu):
    x = x1 if x2 is None else torch.cat([x1, x2], dim=3)
    hi, lo = unpack_conv3x3(packed, x.size(3), bias.numel())
    y = F.conv2d(x.permute(0, 3, 1, 2).double(), (hi + lo).double(), bias.double(), padding=1).permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual.double()
    if relu:
        y = y.clamp_min(0)
    return y.float().contiguous()


def emulated_general(x, mode, stride, dy, dx, ho, wo, packed, bias, relu):
    """x (N, Hi, Wi, Cin) fp32; the documented semantics of mvp_tc_conv_general, tap by tap."""
    n, hi_, wi_, cin = x.shape
    cout = bias.numel()
    g = 4 * cout if mode else cout
    nt = g if g <= 256 else 256
    whi, wlo = unpack_taps(packed, cin, g, len(dy), nt)
    wg = (whi + wlo).double()
    xd = x.double()
    if mode:
        y = torch.einsum('nhwc,gc->nhwg', xd, wg[:, :, 0]).reshape(n, hi_, wi_, 2, 2, cout)      # (.., ky, kx, co)
        y = y.permute(0, 1, 3, 2, 4, 5).reshape(n, 2 * hi_, 2 * wi_, cout) + bias.double()
    else:
        pad = 8
        xp = F.pad(xd, [0, 0, pad, pad + stride * wo, pad, pad + stride * ho])
        y = torch.zeros(n, ho, wo, cout, dtype=torch.float64) + bias.double()
        for t in range(len(dy)):
            ys = pad + dy[t]
            xs = pad + dx[t]
            patch = xp[:, ys:ys + stride * ho:stride, xs:xs + stride * wo:stride, :]
            y = y + torch.einsum('nhwc,gc->nhwg', patch, wg[:, :, t])
    return (y.clamp_min(0) if relu else y).float().contiguous()


def test_pack_round_trip(monkeypatch):
    class _F:
        tc_conv3x3_nt = staticmethod(conv3x3_nt)
        tc_conv3x3_pair_supported = staticmethod(lambda cout, h: h > 8)

    class _E:
        fused_cuda = _F
    monkeypatch.setattr(net2d, 'load_ext', lambda: _E)
    torch.manual_seed(0)
    for cin, cout in [(64, 64), (128, 64), (256, 512), (512, 256)]:
        w = torch.randn(cout, cin, 3, 3)
        packed, b = net2d.pack_conv3x3(w, torch.zeros(cout))
        assert packed.full.numel() == cin * cout * 9 * 4
        hi, lo = unpack_conv3x3(packed.full, cin, cout)
        assert torch.equal(hi, w.bfloat16().float())
        assert (hi + lo - w).abs().max() < 2 ** -16 * w.abs().max()
        # the CTA-pair packing (narrow layers): the same values in blocks of half the width
        assert packed.half is not None
        if packed.half is not None:
            hi2, lo2 = unpack_conv3x3(packed.half, cin, cout, nt=conv3x3_nt(cout) // 2)
            assert torch.equal(hi2, hi) and torch.equal(lo2, lo)


def test_plan_matches_module_features(monkeypatch):
    from mvpnet_b200.unet import UNetResNet34
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        net = UNetResNet34(20, p=0.5, pretrained=False)
    synthetic.fill_parameters(net, seed=4)
    net.eval()

    class _Fused:
        """CPU stand-ins with the extension's signatures: a "planar" tensor is simply the flat fp32 NHWC data."""
        @staticmethod
        def tc_conv3x3_nt(cout):
            return conv3x3_nt(cout)

        @staticmethod
        def split_planar(x):
            return x.reshape(-1).clone()

        @staticmethod
        def merge_planar(p, n, h, w, c):
            return p.reshape(n, h, w, c).clone()

        @staticmethod
        def tc_conv3x3_pair_supported(cout, h):
            return False

        @staticmethod
        def tc_conv3x3(x1, c1, x2, c2, n, h, w, packed, bias, residual, relu, nhwc_out, pair):
            cout = bias.numel()
            y = emulated_conv(x1.reshape(n, h, w, c1), None if x2 is None else x2.reshape(n, h, w, c2), packed, bias,
                              None if residual is None else residual.reshape(n, h, w, cout), relu)
            return y if nhwc_out else y.reshape(-1)

        @staticmethod
        def tc_conv_general(x, cin, n, hi_, wi_, mode, stride, dy, dx, ho, wo, packed, bias, relu):
            return emulated_general(x.reshape(n, hi_, wi_, cin), mode, stride, dy, dx, ho, wo, packed, bias, relu).reshape(-1)

        @staticmethod
        def unfold_stem(img):
            n, _, h, w = img.shape
            out = torch.zeros(n, h, w, 32)
            xp = F.pad(img, [3, 3])
            for kx in range(7):
                out[..., kx * 3:kx * 3 + 3] = xp[:, :, :, kx:kx + w].permute(0, 2, 3, 1)
            return out.reshape(-1)

        @staticmethod
        def maxpool3x3s2_planar(x, n, h, w, c):
            y = F.max_pool2d(x.reshape(n, h, w, c).permute(0, 3, 1, 2), 3, 2, 1)
            return y.permute(0, 2, 3, 1).contiguous().reshape(-1)

    class _Ext:
        fused_cuda = _Fused

    monkeypatch.setattr(net2d, 'load_ext', lambda: _Ext)
    plan = net2d.FastUNetResNet34(net)
    x = torch.randn(2, 3, 24, 40)        # padded to 32 x 48 inside; the deepest level is 2 x 3 pixels
    with torch.no_grad():
        want = net.features(x)
        got = plan.features_nhwc(x).permute(0, 3, 1, 2)
    assert got.shape == want.shape
    assert (got - want).abs().max() <= 2e-5 * want.abs().max()
