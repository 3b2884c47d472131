"""Fused inference path ("fast path") for the MVPNet hot path on B200.

What the reference runs as ~150 small kernels per chunk (pn2ssg.py:87-118, modules.py, mvpnet_3d.py:88-118)
— gather, subtract, concat, 3 x (conv, BN, ReLU), max, each round-tripping a (B, C, M, K) tensor
through HBM — runs here as:

   geometry (depends on xyz only; own stream, overlaps the 2D network)
       4 x [fps -> centroid gather -> ball_query]   +   4 x knn_distance(3-NN)
   features
       feature_aggregation  (pixel gather + relation + MLP + sum)            1 kernel
       4 x set_abstraction   (neighbour gather + centre + MLP + max)         1 kernel each
       4 x feature_propagation (weights + interpolate + concat + MLP)        1 kernel each
       (the segmentation head rides on the last feature_propagation chain)

Tensors between fused kernels are point-major ([B, N, C]); the public (B, C, N) layout of the
reference is restored at the two ends.  BatchNorm is folded into the weights (eval mode only);
training uses the op-by-op modules.  Everything here launches kernels of libmvpnet_b200.so — there
is no eager/CPU fallback inside the fused path; unsupported module configurations fall back to the
op-by-op CUDA composition in modules.py (still this package's kernels).
"""
import contextlib
import os
import weakref

import torch
from torch import nn

from . import load_ext


# ------------------------------------------------------------------------------------------------
# optional per-stage timing (bench.py): CUDA events on the stream the stage is launched on
# ------------------------------------------------------------------------------------------------
class StageTimer:
    """with engine.profile() as t: ...forward...; t.summary() -> {stage: [ms, ...]}"""

    def __init__(self):
        self.records = []

    @contextlib.contextmanager
    def stage(self, name):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        yield
        e.record()
        self.records.append((name, s, e))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, s, e in self.records:
            out.setdefault(name, []).append(s.elapsed_time(e))
        return out


_TIMER = None


@contextlib.contextmanager
def profile():
    global _TIMER
    _TIMER = StageTimer()
    try:
        yield _TIMER
    finally:
        _TIMER = None


def _stage(name):
    return _TIMER.stage(name) if _TIMER is not None else contextlib.nullcontext()


# ------------------------------------------------------------------------------------------------
# weight preparation
# ------------------------------------------------------------------------------------------------
def _round_up(x, m):
    return (x + m - 1) // m * m


def _fold(conv, bn):
    """1x1 conv (+ eval BatchNorm) -> (W [cout, cin], b [cout]) in float64 for a clean fold."""
    w = conv.weight.detach().double().reshape(conv.out_channels, conv.in_channels)
    b = conv.bias.detach().double() if conv.bias is not None else torch.zeros(conv.out_channels, dtype=torch.float64, device=w.device)
    if bn is not None:
        s = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
        w = w * s[:, None]
        b = (b - bn.running_mean.detach().double()) * s + bn.bias.detach().double()
    return w, b


class Chain:
    """A SharedMLP (+ optional extra 1x1 layers) packed for the fused kernels: k-major, zero padded."""

    def __init__(self, layers, cin_true, device):
        # layers: list of (conv, bn_or_None, relu: bool)
        self.wts, self.biases, self.relu = [], [], []
        cin_pad = _round_up(cin_true, 4)
        for conv, bn, relu in layers:
            w, b = _fold(conv, bn)
            cout, cin = w.shape
            cout_pad = _round_up(cout, 8)
            wt = torch.zeros(cin_pad, cout_pad, dtype=torch.float64, device=w.device)
            wt[:cin, :cout] = w.t()
            bias = torch.zeros(cout_pad, dtype=torch.float64, device=w.device)
            bias[:cout] = b
            self.wts.append(wt.float().contiguous().to(device))
            self.biases.append(bias.float().contiguous().to(device))
            self.relu.append(1 if relu else 0)
            cin_pad = cout_pad
        self.out_channels = layers[-1][0].out_channels

    def args(self):
        return self.wts, self.biases, self.relu, self.out_channels


class TcChain:
    """The same chain packed for the tcgen05 kernels (csrc/tc_mlp.cu): every fp32 weight split into
    bf16 hi + bf16 lo, K padded to 16, N padded to 16, stored per 256-wide block of output channels as
    [K/8][nb][8] (the kernel's shared-memory operand order), bias fp32."""

    def __init__(self, layers, cin_true, device):
        self.w_hi, self.w_lo, self.biases, self.ks, self.ns, self.relu = [], [], [], [], [], []
        k_pad = _round_up(cin_true, 16)
        for conv, bn, relu in layers:
            w, b = _fold(conv, bn)
            cout, cin = w.shape
            n_pad = _round_up(cout, 16)
            wp = torch.zeros(n_pad, k_pad, dtype=torch.float64, device=w.device)
            wp[:cout, :cin] = w
            wp = wp.float()
            hi = wp.bfloat16()
            lo = (wp - hi.float()).bfloat16()
            self.w_hi.append(self._arrange(hi, n_pad, k_pad).to(device))
            self.w_lo.append(self._arrange(lo, n_pad, k_pad).to(device))
            bias = torch.zeros(n_pad, dtype=torch.float64, device=w.device)
            bias[:cout] = b
            self.biases.append(bias.float().contiguous().to(device))
            self.ks.append(k_pad)
            self.ns.append(n_pad)
            self.relu.append(1 if relu else 0)
            k_pad = n_pad
        self.out_channels = layers[-1][0].out_channels

    @staticmethod
    def _arrange(x, n_pad, k_pad):
        blocks = []
        for n0 in range(0, n_pad, 256):
            nb = min(256, n_pad - n0)
            blocks.append(x[n0:n0 + nb].reshape(nb, k_pad // 8, 8).permute(1, 0, 2).contiguous().reshape(-1))
        return torch.cat(blocks).contiguous()

    def args(self):
        return self.w_hi, self.w_lo, self.biases, self.ks, self.ns, self.relu, self.out_channels


# 'tc'  : tcgen05 tensor-core kernels where the chain fits (default), fp32 SIMT kernels elsewhere
# 'simt': fp32 SIMT kernels everywhere (cross-check / fallback)
MLP_BACKEND = os.environ.get('MVPNET_B200_MLP', 'tc')
# copy the 2D feature map to channels-last before the pixel gather (coalesced 256-byte rows) instead of gathering
# 64 strided channels per pixel straight from the NCHW output of the 2D network
FA_CHANNELS_LAST_COPY = os.environ.get('MVPNET_B200_FA_CHANNELS_LAST_COPY', '0') == '1'
_MODE_SA, _MODE_FA, _MODE_FP = 0, 1, 2
# second-generation kernels (csrc/tc2_mlp.cu: pre-split inputs, cp.async gather into the swizzled operand, activations in
# tensor memory) where the chain fits them; '0' keeps every chain on csrc/tc_mlp.cu (cross-check)
TC2 = os.environ.get('MVPNET_B200_TC2', '1') == '1'


def split_rows(x):
    """fp32 (..., C) -> bf16 (2, ..., C): hi = bf16(x), lo = bf16(x - hi) — the pre-split row format of the tc2 kernels."""
    hi = x.bfloat16()
    return torch.stack([hi, (x - hi.float()).bfloat16()])


def _tc2_ok(chain, mode, c):
    return TC2 and isinstance(chain, TcChain) and c > 0 and load_ext().fused_cuda.tc2_supported(chain.ks, chain.ns, mode, c)


def _make_chain(layers, cin_true, device, mode, channels_ok):
    """Pick the tensor-core packing when enabled, channel counts are multiples of 8 and the tile fits."""
    if MLP_BACKEND == 'tc' and channels_ok:
        tc = TcChain(layers, cin_true, device)
        if load_ext().fused_cuda.tc_chain_supported(tc.ks, tc.ns, mode):
            return tc
    return Chain(layers, cin_true, device)


def _mlp_layers(shared_mlp):
    return [(m.conv, m.bn, m.relu is not None) for m in shared_mlp]


_CACHE = {}


def _signature(module):
    """What a packed copy of `module`'s weights depends on: identity and version of every parameter and buffer (in-place
    updates bump `_version`, `.data = ...` reassignment changes `data_ptr`), the device, and whether ANY sub-module is in
    training mode.  Walks parameters() / buffers() directly — state_dict() builds 426 prefixed keys per call."""
    sig = [(t.data_ptr(), int(t._version)) for t in module.parameters()]
    sig += [(t.data_ptr(), int(t._version)) for t in module.buffers()]
    return tuple(sig), next(module.parameters()).device, any(m.training for m in module.modules())


def _cached(module, key, build):
    """Per-module cache of packed weights / plans.  Keyed by id(module) AND checked against a weak reference: ids are
    reused after garbage collection, and a stale hit would silently run another model's weights."""
    sig = (key,) + _signature(module)
    hit = _CACHE.get((id(module), key))
    if hit is None or hit[0] != sig or hit[2]() is not module:
        hit = (sig, build(), weakref.ref(module))
        _CACHE[(id(module), key)] = hit
        for k in [k for k, v in _CACHE.items() if v[2]() is None]:      # drop entries of dead modules
            del _CACHE[k]
    return hit[1]


def invalidate(module=None):
    """Drop the packed weights of `module` (or of every module): they are rebuilt on the next fused forward."""
    for k in [k for k in _CACHE if module is None or k[0] == id(module)]:
        del _CACHE[k]


def _require_eval_fp32(module, *tensors):
    if any(m.training for m in module.modules()):      # a sub-module left in train() would get its BatchNorm folded with running statistics
        raise RuntimeError('mvpnet_b200 fused path is inference-only (BatchNorm folded): call .eval() or use forward()')
    for t in tensors:
        if t is not None and (not t.is_cuda or t.dtype != torch.float32):
            raise RuntimeError('mvpnet_b200 fused path needs float32 CUDA tensors')


# ------------------------------------------------------------------------------------------------
# FeatureAggregation
# ------------------------------------------------------------------------------------------------
def feature_aggregation_rows(fa, pix_split, nv, h, w, image_xyz, knn_indices, points, want_f32=False, want_split=True):
    """FeatureAggregation on the PRE-SPLIT pixel rows of the 2D network (net2d.FastUNetResNet34.features_rows:
    bf16 (2, b*nv, hp, wp, c)).  Returns (out fp32 (b, np, c_out) or None, out_split bf16 (2, b, np, c_out) or None),
    or None when the chain does not fit the tc2 kernel (caller falls back to feature_aggregation())."""
    ext = load_ext()
    _require_eval_fp32(fa, image_xyz, points)
    if fa.mlp is None or not fa.use_relation or knn_indices.size(2) > 4:
        return None
    c = pix_split.size(-1)
    chain = _cached(fa, 'fa' + MLP_BACKEND, lambda: _make_chain(_mlp_layers(fa.mlp), c + 4, points.device, _MODE_FA, c % 8 == 0))
    if not _tc2_ok(chain, _MODE_FA, c):
        return None
    b = points.size(0)
    pix = image_xyz.reshape(b, nv * h * w, 3).contiguous()
    pts = points.transpose(1, 2).contiguous()
    with _stage('feature_aggregation'):
        out, sp = ext.fused_cuda.tc2_feature_aggregation(pix_split, nv, h, w, pix, pts, knn_indices.contiguous(), fa.reduction_name == 'sum',
                                                         *chain.args(), want_f32, want_split)
    return (out if want_f32 else None), (sp if want_split else None)


def feature_aggregation(fa, feat2d, image_xyz, knn_indices, points, point_major_out=False):
    """fa: FeatureAggregation module (eval).  feat2d (b, nv, c, h, w) — the 2D network output viewed per
    chunk, any memory format; image_xyz (b, nv, h, w, 3); knn_indices (b, np, k); points (b, 3, np).
    Returns (b, c_out, np) like FeatureAggregation.forward (or (b, np, c_out) when point_major_out).
    Replaces mvpnet_3d.py:100-109 (both group_points) + :37-61."""
    ext = load_ext()
    _require_eval_fp32(fa, feat2d, image_xyz, points)
    if fa.mlp is None or not fa.use_relation or knn_indices.size(2) > 4:
        raise RuntimeError('fused feature_aggregation supports use_relation=True, an MLP and k <= 4')
    b, nv, c, h, w = feat2d.shape
    chain = _cached(fa, 'fa' + MLP_BACKEND, lambda: _make_chain(_mlp_layers(fa.mlp), c + 4, feat2d.device, _MODE_FA, c % 8 == 0))
    pix = image_xyz.reshape(b, nv * h * w, 3).contiguous()
    pts = points.transpose(1, 2).contiguous()
    fn = ext.fused_cuda.tc_feature_aggregation if isinstance(chain, TcChain) else ext.fused_cuda.feature_aggregation
    with _stage('feature_aggregation'):
        out = fn(feat2d, pix, pts, knn_indices.contiguous(), fa.reduction_name == 'sum', *chain.args())
    return out if point_major_out else out.transpose(1, 2).contiguous()


# ------------------------------------------------------------------------------------------------
# PN2SSG
# ------------------------------------------------------------------------------------------------
def _pn2_supported(net):
    for sa in net.sa_modules:
        if sa.num_centroids <= 0 or sa.max_neighbors != 32 or sa.grouper is None:
            return False
        if not (sa.use_xyz or sa.in_channels == 3):
            return False
    for fp in net.fp_modules:
        if fp.interpolator is None:
            return False
    return True


def pn2_geometry(net, xyz_pm):
    """Everything in PN2SSG.forward that depends on coordinates only (pn2ssg.py:96-110 via
    modules.py:100-101, 21, 135): per SA level FPS -> centroids -> ball_query; per FP level the 3-NN."""
    ext = load_ext()
    levels = [xyz_pm]
    nbrs = []
    cur = xyz_pm
    for i, sa in enumerate(net.sa_modules):
        with _stage('fps%d' % (i + 1)):
            idx = ext.fps_cuda.farthest_point_sample(cur, sa.num_centroids)
        new = torch.gather(cur, 1, idx.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        with _stage('ball_query%d' % (i + 1)):
            nbrs.append(ext.ball_query_cuda.ball_query(new, cur, sa.radius, sa.max_neighbors))
        levels.append(new)
        cur = new
    knn = []
    for i in range(len(net.fp_modules)):
        with _stage('knn_distance%d' % (i + 1)):
            knn.append(ext.knn_distance_cuda.knn_distance(levels[-2 - i], levels[-1 - i], 3))
    return {'xyz': levels, 'nbr': nbrs, 'knn': knn}


def _pn2_chains(net, device):
    sa = []
    for m in net.sa_modules:
        sa.append(_make_chain(_mlp_layers(m.mlp), m.in_channels, device, _MODE_SA, (m.in_channels - 3) % 8 == 0))
    fp = []
    for i, m in enumerate(net.fp_modules):
        layers = _mlp_layers(m.mlp)
        if i == len(net.fp_modules) - 1:           # segmentation head rides on the last chain
            layers = layers + _mlp_layers(net.mlp_seg) + [(net.seg_logit, None, False)]
        # interpolated and skip channel counts must both be multiples of 8 for the 8-channel operand units
        prev = net.fp_modules[i - 1].out_channels if i > 0 else net.sa_modules[-1].out_channels
        ok = prev % 8 == 0 and (m.in_channels - prev) % 8 == 0
        fp.append(_make_chain(layers, m.in_channels, device, _MODE_FP, ok))
    return sa, fp


def pn2_features(net, geo, feature_pm, feature_split=None):
    """The feature side of PN2SSG.forward on precomputed geometry.  feature_pm (B, N, C) fp32 or None and / or
    feature_split bf16 (2, B, N, C) (the pre-split rows a tc2 producer wrote).  Returns seg_logit (B, num_classes, N)."""
    ext = load_ext()
    sa_chains, fp_chains = _cached(net, 'pn2' + MLP_BACKEND, lambda: _pn2_chains(net, geo['xyz'][0].device))
    if len(fp_chains[-1].relu) > 6:
        raise RuntimeError('fused PN2SSG: last FP chain + head exceeds 6 layers')
    # which SA levels run on the tc2 kernel: their input channel count (without xyz) must be 64 / 128 / 256
    cin = [(sa.in_channels - 3) if sa.use_xyz else sa.in_channels for sa in net.sa_modules]
    use2 = [_tc2_ok(sa_chains[i], _MODE_SA, cin[i]) and (i > 0 or feature_pm is not None or feature_split is not None)
            for i in range(len(net.sa_modules))]
    feats = [None]
    f, fs = feature_pm, feature_split
    for i, sa in enumerate(net.sa_modules):
        with _stage('set_abstraction%d' % (i + 1)):
            if use2[i]:
                if fs is None:
                    fs = split_rows(f)
                want_split = i + 1 < len(use2) and use2[i + 1]
                f, fs = ext.fused_cuda.tc2_set_abstraction(fs, geo['xyz'][i], geo['xyz'][i + 1], geo['nbr'][i], *sa_chains[i].args(),
                                                           True, want_split)
                fs = fs if want_split else None
            else:
                if f is None and fs is not None:          # only reachable with a caller-supplied split input
                    f = fs[0].float() + fs[1].float()
                fn = ext.fused_cuda.tc_set_abstraction if isinstance(sa_chains[i], TcChain) else ext.fused_cuda.set_abstraction
                f, fs = fn(f, geo['xyz'][i], geo['xyz'][i + 1], geo['nbr'][i], *sa_chains[i].args()), None
        feats.append(f)
    x = feats[-1]
    for i, fp in enumerate(net.fp_modules):
        idx, d2 = geo['knn'][i]
        fn = ext.fused_cuda.tc_feature_propagation if isinstance(fp_chains[i], TcChain) else ext.fused_cuda.feature_propagation
        with _stage('feature_propagation%d' % (i + 1)):
            x = fn(x, idx, d2, feats[-2 - i], fp.interpolator._eps, *fp_chains[i].args())
    return x.transpose(1, 2).contiguous()


def pn2ssg_forward(net, data_batch):
    """Fused PN2SSG.forward (eval).  data_batch: points (B,3,N), optional feature (B,C,N) or
    feature_pm (B,N,C)."""
    points = data_batch['points']
    feature = data_batch.get('feature', None)
    feature_pm = data_batch.get('feature_pm', None)
    _require_eval_fp32(net, points, feature, feature_pm)
    if not _pn2_supported(net):
        return net.forward(data_batch)
    xyz_pm = points.transpose(1, 2).contiguous()
    if feature_pm is None and feature is not None:
        feature_pm = feature.transpose(1, 2).contiguous()
    geo = data_batch.get('geometry', None) or pn2_geometry(net, xyz_pm)
    return {'seg_logit': pn2_features(net, geo, feature_pm)}


# ------------------------------------------------------------------------------------------------
# MVPNet3D
# ------------------------------------------------------------------------------------------------
class _Streams:
    geo = {}          # one side stream per device (a process-global stream would silently serialise a second GPU's forward)


FOLD_NET2D_BN = os.environ.get('MVPNET_B200_FOLD_BN', '1') == '1'


def _folded_net2d(net):
    """Inference copy of the 2D UNet with every eval-mode BatchNorm folded into the preceding (transposed)
    convolution (torch.nn.utils.fusion.fuse_conv_bn_eval): removes ~44 BatchNorm launches (4 ms of the 37 ms
    2D network at 160 views).  The original module and its state_dict are untouched; anything that is not the
    UNetResNet34 of this package is returned as is."""
    from .unet import UNetResNet34
    if not FOLD_NET2D_BN or not isinstance(net, UNetResNet34):
        return net

    def build():
        import copy
        from torch.nn.utils.fusion import fuse_conv_bn_eval
        m = copy.deepcopy(net).eval()
        m.encoder0 = fuse_conv_bn_eval(m.encoder0, m.bn)
        m.bn = nn.Identity()
        for layer in (m.encoder1, m.encoder2, m.encoder3, m.encoder4):
            for blk in layer:
                blk.conv1, blk.bn1 = fuse_conv_bn_eval(blk.conv1, blk.bn1), nn.Identity()
                blk.conv2, blk.bn2 = fuse_conv_bn_eval(blk.conv2, blk.bn2), nn.Identity()
                if blk.downsample is not None:
                    blk.downsample = nn.Sequential(fuse_conv_bn_eval(blk.downsample[0], blk.downsample[1]))
        for name in ('deconv4', 'decoder3', 'deconv3', 'decoder2', 'deconv2', 'decoder1', 'deconv1', 'decoder0'):
            seq = getattr(m, name)
            fused = fuse_conv_bn_eval(seq[0], seq[1], transpose=isinstance(seq[0], nn.ConvTranspose2d))
            setattr(m, name, nn.Sequential(fused, seq[2]))
        return m
    return _cached(net, 'net2d_folded', build)


# 'tc'   : the UNet's 3x3 convolutions on this package's tcgen05 kernel (net2d.py), the rest on cuDNN channels-last
# 'cudnn': the whole 2D network on cuDNN fp32 (BatchNorm folded)
NET2D_BACKEND = os.environ.get('MVPNET_B200_NET2D', 'tc')


def _tc_net2d(net):
    from .unet import UNetResNet34
    if NET2D_BACKEND != 'tc' or not isinstance(net, UNetResNet34):
        return None

    def build():
        from .net2d import FastUNetResNet34
        return FastUNetResNet34(net)
    return _cached(net, 'net2d_tc', build)


def mvpnet3d_forward(model, data_batch, overlap=True):
    """Fused MVPNet3D.forward (eval).  Everything that depends on coordinates only runs on a side stream
    while the 2D network runs on the main stream: the data side (depth unprojection + 2D->3D k-NN, when the
    batch carries `depth`/`pose`/`cam_inv` instead of precomputed `image_xyz`/`knn_indices`) and the geometry of
    the 3D network (FPS, ball queries, 3-NN).  Features then flow 2D net -> feature_aggregation -> PN2SSG chain.

    data_batch: images (b,nv,3,h,w), points (b,3,np) and EITHER image_xyz (b,nv,h,w,3) + knn_indices (b,np,k)
    (the reference's DataLoader output, mvpnet_3d.py:88-109) OR depth (b,nv,h,w) + pose (b,nv,4,4) +
    cam_inv (b,nv,3,3) [+ chunk_box (b,4), k]."""
    if 'images_u8' in data_batch or 'depth_mm' in data_batch:
        data_batch = decode_stored_inputs(data_batch)
    images = data_batch['images']
    points = data_batch['points']
    _require_eval_fp32(model, images, points)
    net3d = model.net_3d
    from_depth = 'knn_indices' not in data_batch
    k = int(data_batch.get('k', 3)) if from_depth else data_batch['knn_indices'].size(2)
    if not _pn2_supported(net3d) or model.feat_aggreg.mlp is None or not model.feat_aggreg.use_relation or k > 4:
        if from_depth:
            data_batch = dict(data_batch)
            data_batch.update(_data_side(data_batch, points.transpose(1, 2).contiguous(), k))
        return model.forward(data_batch)
    b, nv, _, h, w = images.shape
    main = torch.cuda.current_stream()
    xyz_pm = points.transpose(1, 2).contiguous()

    def coordinate_work():
        rg = _data_side(data_batch, xyz_pm, k) if from_depth else data_batch
        return rg, pn2_geometry(net3d, xyz_pm)

    if overlap:
        dev_index = images.device.index if images.device.index is not None else torch.cuda.current_device()
        if dev_index not in _Streams.geo:
            _Streams.geo[dev_index] = torch.cuda.Stream(device=images.device)
        side = _Streams.geo[dev_index]
        side.wait_stream(main)                     # orders reuse of last call's buffers, keeps overlap
        with torch.cuda.stream(side):
            rg, geo = coordinate_work()
    else:
        rg, geo = coordinate_work()
    fa = model.feat_aggreg
    fa_chain = _cached(fa, 'fa' + MLP_BACKEND, lambda: _make_chain(_mlp_layers(fa.mlp), 64 + 4, images.device, _MODE_FA, True)) \
        if fa.in_channels == 64 else None
    with _stage('net_2d'):
        plan = _tc_net2d(model.net_2d)
        rows = plan is not None and fa_chain is not None and _tc2_ok(fa_chain, _MODE_FA, 64)
        if rows:                  # tcgen05 convolutions; the last one writes pre-split pixel rows for the tc2 gather
            pix_split = plan.features_rows(images.reshape(b * nv, *images.shape[2:]))
        elif plan is not None:    # tcgen05 convolutions, fp32 NHWC end to end; (n, c, h, w) view with channel stride 1
            feat2d = plan.features_nhwc(images.reshape(b * nv, *images.shape[2:])).permute(0, 3, 1, 2)
        else:
            feat2d = _folded_net2d(model.net_2d).features(images.reshape(b * nv, *images.shape[2:]))
    if not rows:
        if FA_CHANNELS_LAST_COPY and feat2d.stride(1) != 1:
            with _stage('feat2d_to_channels_last'):
                feat2d = feat2d.contiguous(memory_format=torch.channels_last)   # one pass; pixel rows become 256-byte lines
        feat2d = feat2d.view(b, nv, *feat2d.shape[1:])
    if overlap:
        main.wait_stream(side)
        if not torch.cuda.is_current_stream_capturing():      # (a captured graph owns its private pool: nothing is reused inside it)
            for t in list(geo['xyz']) + list(geo['nbr']) + [x for pair in geo['knn'] for x in pair] + [rg['image_xyz'], rg['knn_indices']]:
                if torch.is_tensor(t):
                    t.record_stream(main)  # allocated on the side stream, consumed on the caller's: keep the allocator from reusing them early
    if rows:
        _, fa_split = feature_aggregation_rows(fa, pix_split, nv, h, w, rg['image_xyz'], rg['knn_indices'], points)
        return {'seg_logit': pn2_features(net3d, geo, None, fa_split)}
    fa_pm = feature_aggregation(fa, feat2d, rg['image_xyz'], rg['knn_indices'], points, point_major_out=True)
    return {'seg_logit': pn2_features(net3d, geo, fa_pm)}


IMAGE_MEAN, IMAGE_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)      # the reference's normaliser (mvpnet/config: ImageNet statistics)


def decode_stored_inputs(data_batch):
    """The formats the dataset stores -> the model's inputs, on the device (the host then uploads 4x / 2x fewer bytes):
    `images_u8` (b, nv, h, w, 3) uint8 -> `images` (b, nv, 3, h, w) float32 = (u8 / 255 - mean) / std
    (scannet_2d3d.py:229-246; `image_mean` / `image_std` in the batch or the ImageNet statistics);
    `depth_mm` (b, nv, h, w) int16 holding uint16 millimetres -> `depth` float32 metres (scannet_2d3d.py:249-251)."""
    ext = load_ext()
    out = dict(data_batch)
    if 'images_u8' in out:
        with _stage('decode_inputs'):
            out['images'] = ext.unproject_cuda.decode_rgb_u8(out.pop('images_u8').contiguous(), list(out.get('image_mean', IMAGE_MEAN)),
                                                             list(out.get('image_std', IMAGE_STD)))
    if 'depth_mm' in out:
        with _stage('decode_inputs'):
            out['depth'] = ext.unproject_cuda.decode_depth_u16(out.pop('depth_mm').contiguous())
    return out


def _data_side(data_batch, xyz_pm, k):
    from .data import unproject_and_knn
    return unproject_and_knn(data_batch['depth'], None, data_batch['pose'], xyz_pm, k=k,
                             chunk_box=data_batch.get('chunk_box'), cam_inv=data_batch['cam_inv'])


# ------------------------------------------------------------------------------------------------
# CUDA-graph replay of the whole forward (static shapes): removes ~800 launches' worth of host work per step
# ------------------------------------------------------------------------------------------------
class GraphedForward:
    """Captures `model.fast_forward(batch)` (both streams, the 2D network included) into one CUDA graph.
    Call with a batch of the same shapes/dtypes: inputs are copied into the captured static buffers, the graph
    is replayed and the static output tensor is returned (valid until the next call)."""

    def __init__(self, model, example_batch, warmup=3):
        self.model = model
        self.static_in = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in example_batch.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                      # cuDNN autotune, lazy inits and chain caches happen here
                model.fast_forward(self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = model.fast_forward(self.static_in)['seg_logit']

    def load(self, batch):
        """Phase 1: copy the batch into the captured static input buffers (the caller's tensors are free afterwards)."""
        for k, v in batch.items():
            if torch.is_tensor(v):
                self.static_in[k].copy_(v, non_blocking=True)

    def replay(self):
        """Phase 2: replay the graph on the loaded inputs; returns the static output tensor."""
        self.graph.replay()
        return self.static_out

    def __call__(self, batch):
        self.load(batch)
        return self.replay()


class PipelinedForward:
    """Throughput API from HOST buffers.  `submit(host_batch)` uploads the batch (pinned host tensors) on a copy stream,
    runs `forward(prepare(device_batch))` and downloads the result into a pinned host buffer on a second copy stream;
    with `depth` >= 2 buffer sets the upload of batch i+1 and the download of batch i-1 overlap the compute of batch i
    (PCIe is full duplex, the copy engines are idle otherwise).  Every submission still performs its own H2D and D2H
    copies; nothing is cached between submissions.

    forward: callable(device_batch) -> result tensor (may be a static CUDA-graph output: it is copied out before the
    next submission of the same slot can overwrite it), or a LIST of `depth` callables, one per buffer set ("lanes"):
    each lane then computes on its own stream, so that consecutive submissions overlap on the GPU as well — the
    under-occupied tail of one forward (small set-abstraction / propagation levels) runs beside the convolutions of
    the next.  Lane callables must not share mutable state (e.g. one GraphedForward per lane).
    Returns (pinned host tensor, event): the tensor holds the result once the event has completed."""

    def __init__(self, forward, example_host_batch, device, prepare=None, depth=2):
        self.lanes = isinstance(forward, (list, tuple))
        if self.lanes:
            depth = len(forward)
        self.forward, self.prepare, self.depth, self.i = forward, prepare, depth, 0
        self.h2d, self.d2h = torch.cuda.Stream(device), torch.cuda.Stream(device)
        self.compute = [torch.cuda.Stream(device) for _ in range(depth)] if self.lanes else None
        self.slots = [{k: torch.empty(v.shape, dtype=v.dtype, device=device) for k, v in example_host_batch.items() if torch.is_tensor(v)}
                      for _ in range(depth)]
        ev = lambda: [torch.cuda.Event() for _ in range(depth)]
        self.in_ready, self.in_free, self.out_ready, self.out_free = ev(), ev(), ev(), ev()
        self.out_dev = [None] * depth
        self.out_host = [None] * depth

    def submit(self, host_batch):
        s = self.i % self.depth
        caller = torch.cuda.current_stream()
        compute = self.compute[s] if self.lanes else caller
        fwd = self.forward[s] if self.lanes else self.forward
        with torch.cuda.stream(self.h2d):
            self.h2d.wait_event(self.in_free[s])          # the forward that last used this buffer set has read it
            for k, dst in self.slots[s].items():
                dst.copy_(host_batch[k], non_blocking=True)
            self.in_ready[s].record(self.h2d)
        with torch.cuda.stream(compute):
            if self.lanes and self.i < self.depth:
                compute.wait_stream(caller)               # first use of the lane: order after the caller's set-up work
            compute.wait_event(self.in_ready[s])
            dev = self.slots[s]
            batch = self.prepare(dev) if self.prepare is not None else dev
            if hasattr(fwd, 'load') and hasattr(fwd, 'replay'):
                # two-phase forwards (GraphedForward) copy the inputs into their own static buffers first: the upload buffers
                # are free as soon as that copy is queued, so the NEXT upload into this slot overlaps this forward instead
                # of waiting for it to end (measured: the lane idled ~0.5 ms per step waiting for its upload)
                fwd.load(batch)
                self.in_free[s].record(compute)
                out = fwd.replay()
            else:
                out = fwd(batch)
                self.in_free[s].record(compute)
            if self.out_dev[s] is None:
                self.out_dev[s] = torch.empty_like(out)
                self.out_host[s] = torch.empty(out.shape, dtype=out.dtype).pin_memory()
            compute.wait_event(self.out_free[s])          # the download that last used this output buffer has left it
            self.out_dev[s].copy_(out)
            self.out_ready[s].record(compute)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(self.out_ready[s])
            self.out_host[s].copy_(self.out_dev[s], non_blocking=True)
            self.out_free[s].record(self.d2h)
        self.i += 1
        return self.out_host[s], self.out_free[s]
