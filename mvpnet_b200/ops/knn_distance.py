"""3-NN with squared distances — mirrors mvpnet/ops/knn_distance.py:5-36."""
from ._util import _NoGrad, channels_last, ext


class KNNDistanceFunction(_NoGrad):
    @staticmethod
    def forward(ctx, query_xyz, key_xyz, k):
        ctx.num_inputs = 3
        index, distance = ext().knn_distance_cuda.knn_distance(query_xyz, key_xyz, k)
        ctx.mark_non_differentiable(index, distance)
        return index, distance


def knn_distance(query, key, k, transpose=True):
    """query (B, 3, N1), key (B, 3, N2), k == 3 -> index int64 (B, N1, 3), squared distance (B, N1, 3),
    ascending; equal distances resolve to the lower key index."""
    return KNNDistanceFunction.apply(channels_last(query, transpose), channels_last(key, transpose), k)
