"""Farthest point sampling — mirrors mvpnet/ops/fps.py:5-31."""
from ._util import _NoGrad, channels_last, ext


class FarthestPointSampleFunction(_NoGrad):
    @staticmethod
    def forward(ctx, points, num_centroids):
        ctx.num_inputs = 2
        index = ext().fps_cuda.farthest_point_sample(points, num_centroids)
        ctx.mark_non_differentiable(index)
        return index


def farthest_point_sample(points, num_centroids, transpose=True):
    """points (B, 3, N) [or (B, N, 3) with transpose=False] -> int64 (B, num_centroids).
    The first centroid is always index 0; ties follow the reference kernel (see csrc/fps.cu)."""
    return FarthestPointSampleFunction.apply(channels_last(points, transpose), num_centroids)
