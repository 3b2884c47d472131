"""feature_interpolate — mirrors mvpnet/ops/interpolate.py:5-34.  Backward: deterministic scatter for float32 (see
group_points.py)."""
import torch

from ._util import deterministic, ext


class FeatureInterpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feature, index, weight):
        ctx.save_for_backward(index, weight)
        ctx.n = feature.size(2)
        return ext().interpolate_cuda.interpolate_forward(feature, index, weight)

    @staticmethod
    def backward(ctx, *grad_out):
        index, weight = ctx.saved_tensors
        g = grad_out[0]
        fn = ext().interpolate_cuda.interpolate_backward_det if deterministic(g) else ext().interpolate_cuda.interpolate_backward
        grad = fn(g, index, weight, ctx.n)
        return grad, None, None


def feature_interpolate(feature, index, weight):
    """feature (B, C, N1), index int64 (B, N2, 3), weight (B, N2, 3) -> (B, C, N2)."""
    return FeatureInterpolate.apply(feature, index, weight)
