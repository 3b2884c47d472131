"""The 2D network of MVPNet3D.forward (mvpnet_3d.py:94-99; architecture unet_resnet34.py:9-125) on this package's
tensor-core convolution (csrc/tc_conv.cu) for inference.

Activations are "split-planar" (bf16 hi + bf16 lo planes of 8-channel slabs, include/mvpnet_b200.h) from the stem to
the last decoder layer.  Every 3x3 / stride-1 convolution (92 % of the network's multiply-adds: ResNet-34's
residual-block convolutions and the four decoder convolutions, whose cat([up, skip]) input is read from the two
tensors in place) runs on csrc/tc_conv.cu; the three stride-2 3x3 convolutions, the 1x1 down-samples, the 2x2
transposed convolutions and the 7x7 stem run on csrc/tc_convg.cu; all as implicit GEMMs on tcgen05 with a bf16
hi/lo split x 3 products (fp32-level accuracy, see tc_mlp.cu).  Eval BatchNorm is folded into weights and bias; ReLU
and the residual add ride in the epilogues.  No cuDNN call and no layout conversion is left.  The final 64-channel
feature map is written fp32 NHWC, the layout the fused FeatureAggregation gather wants (one 256-byte row per pixel).
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


def pack_taps(wg, nt):
    """wg (G, Cin, T) (any float dtype), G % nt == 0, Cin % 16 == 0 -> uint8 tensor in the operand order of the
    tensor-core convolutions: [G/nt][Cin/16][T][hi|lo][2][nt][8] bf16."""
    g, cin, t = wg.shape
    assert g % nt == 0 and cin % 16 == 0
    w = wg.float()
    hi = w.bfloat16()
    lo = (w - hi.float()).bfloat16()
    x = torch.stack([hi, lo])                                             # (hl, G, Cin, T)
    x = x.reshape(2, g // nt, nt, cin // 16, 2, 8, t)                     # (hl, nb, n, c, k8, e, t)
    x = x.permute(1, 3, 6, 0, 4, 2, 5).contiguous()                       # (nb, c, t, hl, k8, n, e)
    return x.view(torch.uint8).reshape(-1).contiguous()


def _nt(cout):
    assert cout % 16 == 0 and (cout <= 256 or cout % 256 == 0)
    return cout if cout <= 256 else 256


class PackedConv3x3:
    """Packed weights of one 3x3 convolution: `full` for mvp_tc_conv3x3 (block width Nt), `half` for the CTA-pair kernel
    mvp_tc_conv3x3_pair (block width Nt / 2; used wherever the image has more than 8 rows)."""
    __slots__ = ('full', 'half', 'cout')

    def __init__(self, full, half, cout):
        self.full, self.half, self.cout = full, half, cout

    def cuda(self):
        return PackedConv3x3(self.full.cuda(), None if self.half is None else self.half.cuda(), self.cout)


def pack_conv3x3(weight, bias):
    """weight (Cout, Cin, 3, 3), bias (Cout) -> (PackedConv3x3, fp32 bias) for mvp_tc_conv3x3[_pair] (tap = ky*3 + kx)."""
    cout, cin = weight.shape[0], weight.shape[1]
    assert weight.shape[2:] == (3, 3)
    fz = load_ext().fused_cuda
    nt = int(fz.tc_conv3x3_nt(cout))              # the kernel's block width is part of the layout
    w9 = weight.reshape(cout, cin, 9)
    half = pack_taps(w9, nt // 2) if fz.tc_conv3x3_pair_supported(cout, 16) else None
    return PackedConv3x3(pack_taps(w9, nt), half, cout), bias.float().contiguous()


def pack_conv_taps(weight, bias):
    """weight (Cout, Cin, kh, kw) -> (packed, bias, dy list, dx list) for mvp_tc_conv_general mode 0, 'same'-style
    padding (kh // 2, kw // 2)."""
    cout, cin, kh, kw = weight.shape
    dy = [ky - kh // 2 for ky in range(kh) for _ in range(kw)]
    dx = [kx - kw // 2 for _ in range(kh) for kx in range(kw)]
    return pack_taps(weight.reshape(cout, cin, kh * kw), _nt(cout)), bias.float().contiguous(), dy, dx


def pack_deconv2x2(weight, bias):
    """ConvTranspose2d weight (Cin, Cout, 2, 2) -> (packed, bias) for mvp_tc_conv_general mode 1:
    GEMM column (2*ky + kx) * Cout + co."""
    cin, cout = weight.shape[0], weight.shape[1]
    assert weight.shape[2:] == (2, 2)
    wg = weight.permute(2, 3, 1, 0).reshape(4 * cout, cin, 1)
    assert cout % 16 == 0 and (4 * cout <= 256 or (4 * cout) % 256 == 0)
    return pack_taps(wg, min(4 * cout, 256)), bias.float().contiguous()


def pack_stem7x7(weight, bias):
    """Conv2d weight (Cout, 3, 7, 7) -> (packed, bias, dy, dx) over the row-unfolded input of mvp_unfold_stem
    (channel kx*3 + c, 21 live of 32): seven row taps."""
    cout = weight.shape[0]
    assert weight.shape[1:] == (3, 7, 7)
    wg = torch.zeros(cout, 32, 7, dtype=weight.dtype, device=weight.device)
    wg[:, :21] = weight.permute(0, 3, 1, 2).reshape(cout, 21, 7)           # (co, kx, c, ky)
    return pack_taps(wg, _nt(cout)), bias.float().contiguous(), [ky - 3 for ky in range(7)], [0] * 7


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


def _stage(name):
    from . import engine          # bench.py's per-stage CUDA-event timers (no-op otherwise)
    return engine._stage(name)


def conv3x3(x1, packed, bias, x2=None, residual=None, relu=True, nhwc_out=0, pair=None):
    """x1 [, x2]: Planar inputs (concatenated along channels); residual: Planar or None.
    Returns a Planar; nhwc_out=1: an fp32 (N, H, W, Cout) tensor; nhwc_out=2: row-split (2, N, H, W, Cout) bf16 (hi, lo planes).
    pair: None = the CTA-pair kernel where it applies (narrow layers), False = always the single-CTA kernel."""
    with _stage('net_2d/conv3x3'):
        return _conv3x3(x1, packed, bias, x2, residual, relu, nhwc_out, pair)


def _conv3x3(x1, packed, bias, x2, residual, relu, nhwc_out, pair=None):
    fz = load_ext().fused_cuda
    pair = pair is not False and packed.half is not None and fz.tc_conv3x3_pair_supported(packed.cout, x1.h)
    out = fz.tc_conv3x3(x1.data, x1.c, None if x2 is None else x2.data, 0 if x2 is None else x2.c, x1.n, x1.h, x1.w,
                        packed.half if pair else packed.full, bias, None if residual is None else residual.data, relu, int(nhwc_out), pair)
    return out if nhwc_out else Planar(out, x1.n, x1.h, x1.w, bias.numel())


def conv_general(x, packed, bias, dy, dx, stride=1, relu=True):
    """taps (dy, dx) convolution with stride 1 or 2 on a Planar (mvp_tc_conv_general mode 0)."""
    ho, wo = (x.h - 1) // stride + 1, (x.w - 1) // stride + 1
    with _stage('net_2d/conv_general'):
        out = load_ext().fused_cuda.tc_conv_general(x.data, x.c, x.n, x.h, x.w, 0, stride, dy, dx, ho, wo, packed, bias, relu)
    return Planar(out, x.n, ho, wo, bias.numel())


def deconv2x2(x, packed, bias, relu=True):
    """2x2 / stride-2 transposed convolution on a Planar (mvp_tc_conv_general mode 1)."""
    with _stage('net_2d/conv_general'):
        out = load_ext().fused_cuda.tc_conv_general(x.data, x.c, x.n, x.h, x.w, 1, 1, [0], [0], 2 * x.h, 2 * x.w, packed, bias, relu)
    return Planar(out, x.n, 2 * x.h, 2 * x.w, bias.numel())


def maxpool3x3s2(x):
    with _stage('net_2d/pool_unfold'):
        out = load_ext().fused_cuda.maxpool3x3s2_planar(x.data, x.n, x.h, x.w, x.c)
    return Planar(out, x.n, (x.h - 1) // 2 + 1, (x.w - 1) // 2 + 1, x.c)


class FastUNetResNet34:
    """Inference plan for a UNetResNet34 (this package's unet.py or the reference's, same attribute names): every
    layer on this package's tensor-core kernels, activations split-planar from the stem to the last decoder layer."""

    def __init__(self, net):
        self.device = next(net.parameters()).device
        e0 = net.encoder0
        if e0.kernel_size != (7, 7) or e0.stride != (1, 1) or e0.padding != (3, 3) or e0.in_channels != 3:
            raise RuntimeError('FastUNetResNet34: unexpected stem geometry')
        self.stem = pack_stem7x7(*fold_conv_bn(e0, net.bn))
        self.layers = []
        for layer in (net.encoder1, net.encoder2, net.encoder3, net.encoder4):
            blocks = []
            for blk in layer:
                e = {'stride': blk.conv1.stride[0]}
                w1, b1 = fold_conv_bn(blk.conv1, blk.bn1)
                e['conv1'] = pack_conv3x3(w1, b1) if e['stride'] == 1 else pack_conv_taps(w1, b1)
                e['conv2'] = pack_conv3x3(*fold_conv_bn(blk.conv2, blk.bn2))
                e['down'] = None
                if blk.downsample is not None:
                    assert blk.downsample[0].stride[0] == e['stride'] and blk.downsample[0].kernel_size == (1, 1)
                    e['down'] = pack_conv_taps(*fold_conv_bn(blk.downsample[0], blk.downsample[1]))
                blocks.append(e)
            self.layers.append(blocks)
        self.dec = []
        for up, fuse in ((net.deconv4, net.decoder3), (net.deconv3, net.decoder2), (net.deconv2, net.decoder1), (net.deconv1, net.decoder0)):
            self.dec.append({'up': pack_deconv2x2(*fold_conv_bn(up[0], up[1])), 'fuse': pack_conv3x3(*fold_conv_bn(fuse[0], fuse[1]))})

    @torch.no_grad()
    def features_rows(self, x):
        """image (n,3,h,w) fp32 -> the 64-channel feature map as PRE-SPLIT pixel rows: bf16 (2, n, hp, wp, 64) = (hi, lo)
        planes over the padded image (hp, wp multiples of 16) — what mvp_tc2_feature_aggregation gathers."""
        return self._run(x, 2)

    @torch.no_grad()
    def features_nhwc(self, x):
        """image (n,3,h,w) fp32 -> 64-channel feature map (n, h, w, 64) fp32, a view of the (padded) NHWC output."""
        n, _, h, w = x.shape
        return self._run(x, 1)[:, :h, :w, :]

    def _run(self, x, out_mode):
        n, _, h, w = x.shape
        pad_h, pad_w = (-h) % 16, (-w) % 16
        if pad_h or pad_w:
            x = F.pad(x, [0, pad_w, 0, pad_h])
        hp, wp = h + pad_h, w + pad_w
        with _stage('net_2d/pool_unfold'):
            x = Planar(load_ext().fused_cuda.unfold_stem(x.contiguous()), n, hp, wp, 32)
        x = conv_general(x, *self.stem, stride=1, relu=True)
        skips = [x]
        x = maxpool3x3s2(x)
        for li, blocks in enumerate(self.layers):
            for e in blocks:
                identity = x
                if e['stride'] == 1:
                    y = conv3x3(x, *e['conv1'], relu=True)
                else:
                    y = conv_general(x, *e['conv1'], stride=e['stride'], relu=True)
                if e['down'] is not None:
                    identity = conv_general(x, *e['down'], stride=e['stride'], relu=False)
                x = conv3x3(y, *e['conv2'], residual=identity, relu=True)
            if li < 3:
                skips.append(x)
        for i, (d, skip) in enumerate(zip(self.dec, (skips[3], skips[2], skips[1], skips[0]))):
            x = conv3x3(deconv2x2(x, *d['up'], relu=True), *d['fuse'], x2=skip, relu=True, nhwc_out=(out_mode if i == 3 else 0))
        return x
