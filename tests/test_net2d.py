"""Host logic of the tensor-core 2D network plan (mvpnet_b200/net2d.py) on CPU: weight packing round trip and the
wiring of the plan (BatchNorm folding, residuals, concat-free decoder) against UNetResNet34.features, with the CUDA
convolution replaced by a torch emulation that consumes the PACKED weights."""
import warnings

import torch
import torch.nn.functional as F

from mvpnet_b200 import net2d, synthetic


def unpack_conv3x3(packed, cin, cout):
    nt = cout if cout <= 256 else 256
    x = packed.view(torch.bfloat16).reshape(cout // nt, cin // 16, 3, 3, 2, 2, nt, 8)   # (nb, c, ky, kx, hl, k8, n, e)
    x = x.permute(4, 0, 6, 1, 5, 7, 2, 3).reshape(2, cout, cin, 3, 3).float()
    return x[0], x[1]


def emulated_conv(x1, x2, packed, bias, residual, relu):
    x = x1 if x2 is None else torch.cat([x1, x2], dim=3)
    hi, lo = unpack_conv3x3(packed, x.size(3), bias.numel())
    y = F.conv2d(x.permute(0, 3, 1, 2).double(), (hi + lo).double(), bias.double(), padding=1).permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual.double()
    if relu:
        y = y.clamp_min(0)
    return y.float().contiguous()


def test_pack_round_trip():
    torch.manual_seed(0)
    for cin, cout in [(64, 64), (128, 64), (256, 512), (512, 256)]:
        w = torch.randn(cout, cin, 3, 3)
        packed, b = net2d.pack_conv3x3(w, torch.zeros(cout))
        assert packed.numel() == cin * cout * 9 * 4
        hi, lo = unpack_conv3x3(packed, cin, cout)
        assert torch.equal(hi, w.bfloat16().float())
        assert (hi + lo - w).abs().max() < 2 ** -16 * w.abs().max()


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
        def split_planar(x):
            return x.reshape(-1).clone()

        @staticmethod
        def merge_planar(p, n, h, w, c):
            return p.reshape(n, h, w, c).clone()

        @staticmethod
        def tc_conv3x3(x1, c1, x2, c2, n, h, w, packed, bias, residual, relu, nhwc_out):
            cout = bias.numel()
            y = emulated_conv(x1.reshape(n, h, w, c1), None if x2 is None else x2.reshape(n, h, w, c2), packed, bias,
                              None if residual is None else residual.reshape(n, h, w, cout), relu)
            return y if nhwc_out else y.reshape(-1)

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
