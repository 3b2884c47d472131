"""Upper face of the boundary: the functions and autograd.Function classes of `mvpnet.ops`
(reference: mvpnet/ops/{fps,ball_query,group_points,knn_distance,interpolate}.py), same names,
argument order and layouts, running on the sm_100a extension."""
from .fps import farthest_point_sample, FarthestPointSampleFunction
from .ball_query import ball_query, ball_query_distance, BallQueryFunction, BallQueryDistanceFunction
from .group_points import group_points, GroupPointsFunction
from .knn_distance import knn_distance, KNNDistanceFunction
from .interpolate import feature_interpolate, FeatureInterpolate

__all__ = ['farthest_point_sample', 'ball_query', 'ball_query_distance', 'group_points', 'knn_distance',
           'feature_interpolate', 'FarthestPointSampleFunction', 'BallQueryFunction',
           'BallQueryDistanceFunction', 'GroupPointsFunction', 'KNNDistanceFunction', 'FeatureInterpolate']
