"""Ball query — mirrors mvpnet/ops/ball_query.py:6-45."""
from ._util import _NoGrad, channels_last, ext


class BallQueryFunction(_NoGrad):
    @staticmethod
    def forward(ctx, query, key, radius, max_neighbors):
        ctx.num_inputs = 4
        index = ext().ball_query_cuda.ball_query(query, key, radius, max_neighbors)
        ctx.mark_non_differentiable(index)
        return index


class BallQueryDistanceFunction(_NoGrad):
    @staticmethod
    def forward(ctx, query, key, radius, max_neighbors):
        ctx.num_inputs = 4
        index, distance = ext().ball_query_distance_cuda.ball_query_distance(query, key, radius, max_neighbors)
        ctx.mark_non_differentiable(index, distance)
        return index, distance


def ball_query(query, key, radius, max_neighbors, transpose=True):
    """query (B, 3, N1), key (B, 3, N2) -> int64 (B, N1, max_neighbors): the first neighbours in key
    order inside the open ball; short rows are padded with their first hit, empty rows are -1."""
    return BallQueryFunction.apply(channels_last(query, transpose), channels_last(key, transpose),
                                   radius, max_neighbors)


def ball_query_distance(query, key, radius, max_neighbors, transpose=True):
    """As ball_query, plus the squared distances (padding slots hold -1)."""
    return BallQueryDistanceFunction.apply(channels_last(query, transpose), channels_last(key, transpose),
                                           radius, max_neighbors)
