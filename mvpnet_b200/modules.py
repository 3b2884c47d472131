"""nn.Module mirror of the reference's hot-path modules, same constructor signatures, forward
signatures and state_dict keys, running on the sm_100a ops:

  SharedMLP / SharedMLPDO / Conv{1,2}dBNReLU   common/nn/modules/{mlp,conv}.py
  batch_index_select                           common/nn/functional.py:125-146
  QueryGrouper, SetAbstraction,
  FeatureInterpolator, FeaturePropagation      mvpnet/models/pn2/modules.py:13-186
  PN2SSG                                       mvpnet/models/pn2/pn2ssg.py:22-118
  FeatureAggregation, MVPNet3D                 mvpnet/models/mvpnet_3d.py:9-118

The forward()s here are the op-by-op composition (training and the parity baseline); the fused
inference path lives in mvpnet_b200/engine.py and is selected with `MVPNet3D.fast_forward` /
`PN2SSG.fast_forward`.
"""
import torch
from torch import nn
import torch.nn.functional as F

from .ops import (ball_query, farthest_point_sample, feature_interpolate, group_points, knn_distance)


# --------------------------------------------------------------------------------------------- nn
class _ConvBNReLU(nn.Module):
    """1x1-style conv -> (BatchNorm) -> (ReLU).  Conv bias only when there is no BN
    (conv.py:16,41).  Children are named conv / bn / relu so state_dict keys match."""
    _conv, _bn = None, None

    def __init__(self, in_channels, out_channels, kernel_size, relu=True, bn=True, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.conv = self._conv(in_channels, out_channels, kernel_size, bias=(not bn), **kwargs)
        self.bn = self._bn(out_channels) if bn else None
        self.relu = nn.ReLU(inplace=True) if relu else None

    def forward(self, x):
        x = self.conv(x)
        if self.bn is not None:
            x = self.bn(x)
        if self.relu is not None:
            x = self.relu(x)
        return x


class Conv1dBNReLU(_ConvBNReLU):
    _conv, _bn = nn.Conv1d, nn.BatchNorm1d


class Conv2dBNReLU(_ConvBNReLU):
    _conv, _bn = nn.Conv2d, nn.BatchNorm2d


class SharedMLP(nn.ModuleList):
    """Pointwise MLP shared over 1 or 2 trailing axes (mlp.py:38-75)."""

    def __init__(self, in_channels, mlp_channels, ndim=1, bn=True):
        super().__init__()
        if ndim not in (1, 2):
            raise ValueError('SharedMLP only supports ndim=(1, 2).')
        self.in_channels, self.out_channels, self.ndim = in_channels, mlp_channels[-1], ndim
        layer = Conv1dBNReLU if ndim == 1 else Conv2dBNReLU
        c = in_channels
        for c_out in mlp_channels:
            self.append(layer(c, c_out, 1, relu=True, bn=bn))
            c = c_out

    def forward(self, x):
        for layer in self:
            x = layer(x)
        return x


class SharedMLPDO(SharedMLP):
    """SharedMLP with dropout after every layer (mlp.py:78-95)."""

    def __init__(self, *args, p=0.5, **kwargs):
        super().__init__(*args, **kwargs)
        self.p = p

    def forward(self, x):
        drop = F.dropout if self.ndim == 1 else F.dropout2d
        for layer in self:
            x = drop(layer(x), p=self.p, training=self.training, inplace=False)
        return x

    def extra_repr(self):
        return 'p={}'.format(self.p)


def batch_index_select(input, index, dim):
    """input (b, ...), index (b, n): per-batch index_select along `dim` (functional.py:125-146)."""
    assert index.dim() == 2, 'Index should be 2-dim.'
    assert input.size(0) == index.size(0), 'Mismatched batch size: {} vs {}'.format(input.size(0), index.size(0))
    shape = [1] * input.dim()
    shape[0], shape[dim] = index.size(0), index.size(1)
    target = list(input.shape)
    target[dim] = -1
    return torch.gather(input, dim, index.view(shape).expand(target))


def _xavier(module):
    for m in module.modules():
        if isinstance(m, (nn.Conv1d, nn.Conv2d, nn.Linear)):
            if m.weight is not None:
                nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)


# -------------------------------------------------------------------------------------------- pn2
class QueryGrouper(nn.Module):
    """ball_query -> group xyz (centred on the centroid) and features (modules.py:13-41).
    Output channel order: features first, then the 3 relative xyz channels."""

    def __init__(self, radius, max_neighbors):
        super().__init__()
        assert radius > 0.0 and max_neighbors > 0
        self.radius, self.max_neighbors = radius, max_neighbors

    def forward(self, new_xyz, xyz, feature, use_xyz):
        with torch.no_grad():
            index = ball_query(new_xyz, xyz, self.radius, self.max_neighbors)
        group_xyz = group_points(xyz, index)
        group_xyz = group_xyz - new_xyz.unsqueeze(-1)
        if feature is None:
            return group_xyz, group_xyz
        group_feature = group_points(feature, index)
        if use_xyz:
            group_feature = torch.cat([group_feature, group_xyz], dim=1)
        return group_feature, group_xyz

    def extra_repr(self):
        return 'radius={}, max_neighbors={}'.format(self.radius, self.max_neighbors)


class SetAbstraction(nn.Module):
    """FPS -> centroids -> QueryGrouper -> SharedMLP(2D) -> max over neighbours (modules.py:44-113).
    num_centroids == 0: one global group around the origin (xyz NOT centred); == -1: no sampling."""

    def __init__(self, in_channels, mlp_channels, num_centroids, radius, max_neighbors, use_xyz):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = mlp_channels[-1]
        self.num_centroids, self.radius, self.max_neighbors, self.use_xyz = num_centroids, radius, max_neighbors, use_xyz
        if use_xyz or in_channels == 0:
            self.in_channels += 3
        self.mlp = SharedMLP(self.in_channels, mlp_channels, ndim=2, bn=True)
        self.grouper = None if num_centroids == 0 else QueryGrouper(radius, max_neighbors)

    def forward(self, xyz, feature=None):
        if self.num_centroids == 0:
            assert feature is not None
            new_xyz = xyz.new_zeros([xyz.size(0), 3, 1])
            group_feature = feature.unsqueeze(2)
            if self.use_xyz:
                group_feature = torch.cat([group_feature, xyz.unsqueeze(2)], dim=1)
        else:
            if self.num_centroids == -1:
                new_xyz = xyz
            else:
                with torch.no_grad():
                    index = farthest_point_sample(xyz, self.num_centroids)
                new_xyz = batch_index_select(xyz, index, dim=2)
            group_feature, _ = self.grouper(new_xyz, xyz, feature, use_xyz=self.use_xyz)
        new_feature = self.mlp(group_feature)
        new_feature, _ = torch.max(new_feature, dim=3)
        return new_xyz, new_feature

    def extra_repr(self):
        return 'num_centroids={}, radius={}, max_neighbors={}, use_xyz={}'.format(
            self.num_centroids, self.radius, self.max_neighbors, self.use_xyz)


class FeatureInterpolator(nn.Module):
    """3-NN inverse-(squared-)distance interpolation of key features at the query points, concatenated
    in front of the query's own features (modules.py:116-153)."""

    def __init__(self, num_neighbors, eps=1e-10):
        super().__init__()
        self.num_neighbors, self._eps = num_neighbors, eps

    def forward(self, query_xyz, key_xyz, query_feature, key_feature):
        with torch.no_grad():
            index, distance = knn_distance(query_xyz, key_xyz, self.num_neighbors)
            inv = 1.0 / torch.clamp(distance, min=self._eps)
            weight = inv / torch.sum(inv, dim=2, keepdim=True)
        out = feature_interpolate(key_feature, index, weight)
        if query_feature is not None:
            out = torch.cat([out, query_feature], dim=1)
        return out

    def extra_repr(self):
        return 'num_neighbors={}'.format(self.num_neighbors)


class FeaturePropagation(nn.Module):
    """FeatureInterpolator -> SharedMLP(1D) (modules.py:156-186); num_neighbors == 0 broadcasts a
    single global feature instead."""

    def __init__(self, in_channels, in_channels_prev, mlp_channels, num_neighbors):
        super().__init__()
        self.in_channels = in_channels + in_channels_prev
        self.out_channels = mlp_channels[-1]
        self.mlp = SharedMLP(self.in_channels, mlp_channels, ndim=1, bn=True)
        if num_neighbors == 0:
            self.interpolator = None
        elif num_neighbors == 3:
            self.interpolator = FeatureInterpolator(num_neighbors)
        else:
            raise ValueError('Expected value 3, but {} given.'.format(num_neighbors))

    def forward(self, dense_xyz, sparse_xyz, dense_feature, sparse_feature):
        if self.interpolator is None:
            assert sparse_xyz.size(2) == 1 and sparse_feature.size(2) == 1
            x = torch.cat([sparse_feature.expand(-1, -1, dense_xyz.size(2)), dense_feature], dim=1)
        else:
            x = self.interpolator(dense_xyz, sparse_xyz, dense_feature, sparse_feature)
        return self.mlp(x)


class PN2SSG(nn.Module):
    """PointNet++ single-scale-grouping segmentation network (pn2ssg.py:22-118)."""

    def __init__(self, in_channels, num_classes,
                 sa_channels=((32, 32, 64), (64, 64, 128), (128, 128, 256), (256, 256, 512)),
                 num_centroids=(2048, 512, 128, 32), radius=(0.1, 0.2, 0.4, 0.8),
                 max_neighbors=(32, 32, 32, 32),
                 fp_channels=((256, 256), (256, 256), (256, 128), (128, 128, 128)),
                 fp_neighbors=(3, 3, 3, 3), seg_channels=(128,), dropout_prob=0.5, use_xyz=True):
        super().__init__()
        self.in_channels, self.num_classes, self.use_xyz = in_channels, num_classes, use_xyz
        n_sa = len(sa_channels)
        assert len(num_centroids) == n_sa and len(radius) == n_sa and len(max_neighbors) == n_sa
        assert len(fp_channels) == n_sa and len(fp_neighbors) == n_sa

        self.sa_modules = nn.ModuleList()
        c = in_channels
        for i in range(n_sa):
            self.sa_modules.append(SetAbstraction(c, sa_channels[i], num_centroids[i], radius[i], max_neighbors[i], use_xyz))
            c = sa_channels[i][-1]
        # skip features: level 0 (the network input) is NOT used as a skip (pn2ssg.py:64-67, 94)
        skip = [0] + [ch[-1] for ch in sa_channels]
        self.fp_modules = nn.ModuleList()
        c = skip[-1]
        for i in range(n_sa):
            self.fp_modules.append(FeaturePropagation(c, skip[-2 - i], fp_channels[i], fp_neighbors[i]))
            c = fp_channels[i][-1]
        self.mlp_seg = SharedMLPDO(c, seg_channels, ndim=1, bn=True, p=dropout_prob)
        self.seg_logit = nn.Conv1d(seg_channels[-1], num_classes, 1, bias=True)
        self.reset_parameters()

    def reset_parameters(self):
        _xavier(self)

    def forward(self, data_batch):
        xyz = data_batch['points']
        feature = data_batch.get('feature', None)
        xyzs, feats = [xyz], [None]
        for sa in self.sa_modules:
            xyz, feature = sa(xyz, feature)
            xyzs.append(xyz)
            feats.append(feature)
        x = feats[-1]
        for i, fp in enumerate(self.fp_modules):
            x = fp(xyzs[-2 - i], xyzs[-1 - i], feats[-2 - i], x)
        x = self.mlp_seg(x)
        return {'seg_logit': self.seg_logit(x)}

    def fast_forward(self, data_batch):
        """Inference through the fused sm_100a kernels (mvpnet_b200/engine.py); same result contract."""
        from . import engine
        return engine.pn2ssg_forward(self, data_batch)


# ------------------------------------------------------------------------------------------- mvpnet
class FeatureAggregation(nn.Module):
    """Per point: k neighbouring pixels' features + relation (dxyz, |dxyz|^2) -> SharedMLP(2D) ->
    sum / max over k (mvpnet_3d.py:9-67).  Channel order: features first, then dx, dy, dz, d2."""

    def __init__(self, in_channels, mlp_channels=(64, 64, 64), reduction='sum', use_relation=True):
        super().__init__()
        self.in_channels, self.use_relation = in_channels, use_relation
        if mlp_channels:
            self.out_channels = mlp_channels[-1]
            self.mlp = SharedMLP(in_channels + (4 if use_relation else 0), mlp_channels, ndim=2, bn=True)
        else:
            self.out_channels = in_channels
            self.mlp = None
        if reduction not in ('sum', 'max'):
            raise ValueError('reduction must be sum or max')
        self.reduction_name = reduction
        self.reset_parameters()

    def reduction(self, x, dim):
        return torch.sum(x, dim) if self.reduction_name == 'sum' else torch.max(x, dim)[0]

    def forward(self, src_xyz, tgt_xyz, feature):
        if self.mlp is None:
            return self.reduction(feature, 3)
        x = feature
        if self.use_relation:
            diff = src_xyz - tgt_xyz.unsqueeze(-1)
            dist = torch.sum(diff ** 2, dim=1, keepdim=True)
            x = torch.cat([feature, diff, dist], dim=1)
        return self.reduction(self.mlp(x), 3)

    def reset_parameters(self):
        _xavier(self)


class MVPNet3D(nn.Module):
    """2D network -> gather k pixel features / pixel xyz per point -> FeatureAggregation -> 3D network
    (mvpnet_3d.py:70-118).  data_batch keys: images (b,nv,3,h,w), image_xyz (b,nv,h,w,3),
    knn_indices (b,np,k) flat pixel ids, points (b,3,np)."""

    def __init__(self, net_2d, net_2d_ckpt_path, net_3d, **feat_aggr_kwargs):
        super().__init__()
        self.net_2d = net_2d
        if net_2d_ckpt_path:
            checkpoint = torch.load(net_2d_ckpt_path, map_location=torch.device('cpu'))
            self.net_2d.load_state_dict(checkpoint['model'])
        self.feat_aggreg = FeatureAggregation(**feat_aggr_kwargs)
        self.net_3d = net_3d

    def forward(self, data_batch):
        images = data_batch['images']
        b, nv, _, h, w = images.size()
        feature_2d = self.net_2d({'image': images.reshape([-1] + list(images.shape[2:]))})['feature']
        knn_indices = data_batch['knn_indices']
        feature_2d = feature_2d.reshape(b, nv, -1, h, w).transpose(1, 2).contiguous().reshape(b, -1, nv * h * w)
        feature_2d = group_points(feature_2d, knn_indices)
        with torch.no_grad():
            image_xyz = data_batch['image_xyz'].permute(0, 4, 1, 2, 3).reshape(b, 3, nv * h * w)
            image_xyz = group_points(image_xyz, knn_indices)
        points = data_batch['points']
        feature_2d3d = self.feat_aggreg(image_xyz, points, feature_2d)
        return self.net_3d({'points': points, 'feature': feature_2d3d})

    def fast_forward(self, data_batch, overlap=True):
        """Inference through the fused sm_100a kernels (mvpnet_b200/engine.py); same result contract."""
        from . import engine
        return engine.mvpnet3d_forward(self, data_batch, overlap=overlap)


class MVPNet2D(nn.Module):
    """2D network only: per-pixel logits gathered at the k nearest pixels of every point and averaged
    (mvpnet/models/mvpnet_2d.py:7-34).  data_batch keys: images (b,nv,3,h,w), knn_indices (b,np,k)."""

    def __init__(self, net_2d):
        super().__init__()
        self.net_2d = net_2d

    def forward(self, data_batch):
        images = data_batch['images']
        b, nv, _, h, w = images.size()
        seg_logit_2d = self.net_2d({'image': images.reshape([-1] + list(images.shape[2:]))})['seg_logit']
        knn_indices = data_batch['knn_indices']
        seg_logit = seg_logit_2d.reshape(b, nv, -1, h, w).transpose(1, 2).contiguous().reshape(b, -1, nv * h * w)
        return {'seg_logit': group_points(seg_logit, knn_indices).mean(-1)}

    def fast_forward(self, data_batch):
        """Inference on the tensor-core 2D network (net2d.py).  The 1x1 logit head is linear, so it is applied AFTER
        the gather + mean over the k pixels: 64-channel rows of the NHWC feature map are gathered (one 256-byte row
        per pixel) instead of materialising per-pixel logits for every view."""
        from . import engine
        images, knn = data_batch['images'], data_batch['knn_indices']
        engine._require_eval_fp32(self, images)
        plan = engine._tc_net2d(self.net_2d)
        if plan is None:
            return self.forward(data_batch)
        b, nv, _, h, w = images.size()
        feat = plan.features_nhwc(images.reshape(b * nv, *images.shape[2:]))          # (b*nv, h, w, c) view of padded rows
        c = feat.size(3)
        pix = knn.reshape(b, -1)                                                      # flat pixel ids v*h*w + y*w + x
        v, rem = pix // (h * w), pix % (h * w)
        rows = feat[(torch.arange(b, device=pix.device).unsqueeze(1) * nv + v), rem // w, rem % w]   # (b, np*k, c)
        mean = rows.reshape(b, knn.size(1), knn.size(2), c).mean(dim=2)                # (b, np, c)
        head = self.net_2d.logit
        logit = torch.matmul(mean, head.weight.reshape(head.out_channels, c).t()) + head.bias
        return {'seg_logit': logit.transpose(1, 2).contiguous()}
