"""The 2D network of MVPNet3D.forward (mvpnet_3d.py:94-99; architecture unet_resnet34.py:9-125) on this package's
tensor-core convolution (csrc/tc_conv.cu) for inference.

Activations are fp32 NHWC end to end.  Every 3x3 / stride-1 convolution (92 % of the network's multiply-adds: all of
ResNet-34's residual-block convolutions except the three strided ones, and the four decoder convolutions, whose
cat([up, skip]) input is read from the two tensors in place) runs as an implicit GEMM on tcgen05 with a bf16 hi/lo
split x 3 products (fp32-level accuracy, see tc_mlp.cu); eval BatchNorm is folded into weights and bias, ReLU and the
residual add ride in the epilogue.  The remaining layers (7x7 stem, max-pool, three stride-2 3x3, three 1x1
down-samples, four 2x2 transposed convolutions) stay on cuDNN / ATen in channels-last fp32 so that no layout
conversion happens anywhere.  The final 64-channel feature map is returned NHWC, which is exactly the layout the
fused FeatureAggregation gather wants (one 256-byte row per pixel).
"""
import torch
from torch import nn
import torch.nn.functional as F

from . import load_ext


def fold_conv_bn(conv, bn):
    """conv (+ eval BatchNorm) -> (weight, bias) in float64."""
    w = conv.weight.detach().double()
    b = conv.bias.detach().double() if conv.bias is not None else torch.zeros(w.size(0 if not isinstance(conv, nn.ConvTranspose2d) else 1),
                                                                              dtype=torch.float64, device=w.device)
    if bn is not None:
        s = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
        if isinstance(conv, nn.ConvTranspose2d):
            w = w * s[None, :, None, None]
        else:
            w = w * s[:, None, None, None]
        b = (b - bn.running_mean.detach().double()) * s + bn.bias.detach().double()
    return w, b


def pack_conv3x3(weight, bias):
    """weight (Cout, Cin, 3, 3), bias (Cout) (any float dtype) -> (packed uint8 tensor, fp32 bias) in the operand order
    of mvp_tc_conv3x3: [Cout/Nt][Cin/16][tap][hi|lo][2][Nt][8] bf16."""
    cout, cin = weight.shape[0], weight.shape[1]
    assert weight.shape[2:] == (3, 3) and cin % 16 == 0 and cout % 16 == 0
    nt = cout if cout <= 256 else 256
    assert cout % nt == 0
    w = weight.float()
    hi = w.bfloat16()
    lo = (w - hi.float()).bfloat16()
    x = torch.stack([hi, lo])                                             # (hl, Cout, Cin, ky, kx)
    x = x.reshape(2, cout // nt, nt, cin // 16, 2, 8, 3, 3)               # (hl, nb, n, c, k8, e, ky, kx)
    x = x.permute(1, 3, 6, 7, 0, 4, 2, 5).contiguous()                    # (nb, c, ky, kx, hl, k8, n, e)
    return x.view(torch.uint8).reshape(-1).contiguous(), bias.float().contiguous()


class Planar:
    """A split-planar activation tensor (see include/mvpnet_b200.h): flat bf16 storage + logical (N, H, W, C)."""
    __slots__ = ('data', 'n', 'h', 'w', 'c')

    def __init__(self, data, n, h, w, c):
        self.data, self.n, self.h, self.w, self.c = data, n, h, w, c

    @staticmethod
    def from_nhwc(x):
        """fp32 (N, H, W, C) contiguous -> split-planar."""
        n, h, w, c = x.shape
        return Planar(load_ext().fused_cuda.split_planar(x), n, h, w, c)

    def to_nhwc(self):
        return load_ext().fused_cuda.merge_planar(self.data, self.n, self.h, self.w, self.c)


def conv3x3(x1, packed, bias, x2=None, residual=None, relu=True, nhwc_out=False):
    """x1 [, x2]: Planar inputs (concatenated along channels); residual: Planar or None.
    Returns a Planar, or an fp32 (N, H, W, Cout) tensor when nhwc_out."""
    out = load_ext().fused_cuda.tc_conv3x3(x1.data, x1.c, None if x2 is None else x2.data, 0 if x2 is None else x2.c,
                                           x1.n, x1.h, x1.w, packed, bias, None if residual is None else residual.data, relu, nhwc_out)
    return out if nhwc_out else Planar(out, x1.n, x1.h, x1.w, bias.numel())


def _nhwc(t):
    """logical NCHW tensor -> contiguous NHWC (no copy when the memory format already is channels-last)."""
    return t.permute(0, 2, 3, 1).contiguous()


def _nchw_view(t):
    """contiguous NHWC tensor -> logical NCHW view in channels-last memory format (no copy)."""
    return t.permute(0, 3, 1, 2)


class FastUNetResNet34:
    """Inference plan for a UNetResNet34 (this package's unet.py or the reference's, same attribute names)."""

    def __init__(self, net):
        dev = next(net.parameters()).device
        self.device = dev
        cl = torch.channels_last

        def cudnn_conv(conv, bn):
            w, b = fold_conv_bn(conv, bn)
            return w.float().contiguous(memory_format=cl), b.float().contiguous()

        def mine(conv, bn):
            w, b = fold_conv_bn(conv, bn)
            return pack_conv3x3(w, b)

        self.stem = cudnn_conv(net.encoder0, net.bn)
        self.stem_stride, self.stem_pad = net.encoder0.stride, net.encoder0.padding
        self.layers = []
        for layer in (net.encoder1, net.encoder2, net.encoder3, net.encoder4):
            blocks = []
            for blk in layer:
                e = {'stride': blk.conv1.stride[0]}
                e['conv1'] = mine(blk.conv1, blk.bn1) if e['stride'] == 1 else cudnn_conv(blk.conv1, blk.bn1)
                e['conv2'] = mine(blk.conv2, blk.bn2)
                e['down'] = cudnn_conv(blk.downsample[0], blk.downsample[1]) if blk.downsample is not None else None
                e['down_stride'] = blk.downsample[0].stride if blk.downsample is not None else None
                blocks.append(e)
            self.layers.append(blocks)
        self.dec = []
        for up, fuse in ((net.deconv4, net.decoder3), (net.deconv3, net.decoder2), (net.deconv2, net.decoder1), (net.deconv1, net.decoder0)):
            wu, bu = fold_conv_bn(up[0], up[1])
            self.dec.append({'up_w': wu.float().contiguous(memory_format=cl), 'up_b': bu.float().contiguous(),
                             'fuse': mine(fuse[0], fuse[1])})

    @torch.no_grad()
    def features_nhwc(self, x):
        """image (n,3,h,w) fp32 -> 64-channel feature map (n, h, w, 64) fp32, a view of the (padded) NHWC output."""
        h, w = x.shape[2], x.shape[3]
        pad_h, pad_w = (-h) % 16, (-w) % 16
        if pad_h or pad_w:
            x = F.pad(x, [0, pad_w, 0, pad_h])
        x = x.contiguous(memory_format=torch.channels_last)
        x = F.relu_(F.conv2d(x, self.stem[0], self.stem[1], self.stem_stride, self.stem_pad))
        skips = [Planar.from_nhwc(_nhwc(x))]
        x = Planar.from_nhwc(_nhwc(F.max_pool2d(x, kernel_size=3, stride=2, padding=1)))
        for li, blocks in enumerate(self.layers):
            for e in blocks:
                identity = x
                if e['stride'] == 1:
                    y = conv3x3(x, *e['conv1'], relu=True)
                else:                      # strided block: cuDNN on the merged fp32 view, results split again
                    xf = _nchw_view(x.to_nhwc())
                    y = Planar.from_nhwc(_nhwc(F.relu_(F.conv2d(xf, e['conv1'][0], e['conv1'][1], e['stride'], 1))))
                    identity = Planar.from_nhwc(_nhwc(F.conv2d(xf, e['down'][0], e['down'][1], e['down_stride'], 0)))
                x = conv3x3(y, *e['conv2'], residual=identity, relu=True)
            if li < 3:
                skips.append(x)
        for i, (d, skip) in enumerate(zip(self.dec, (skips[3], skips[2], skips[1], skips[0]))):
            up = F.relu_(F.conv_transpose2d(_nchw_view(x.to_nhwc()), d['up_w'], d['up_b'], stride=2))
            x = conv3x3(Planar.from_nhwc(_nhwc(up)), *d['fuse'], x2=skip, relu=True, nhwc_out=(i == 3))
        return x[:, :h, :w, :]
