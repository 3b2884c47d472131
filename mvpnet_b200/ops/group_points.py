"""group_points — mirrors mvpnet/ops/group_points.py:5-31.  The backward is the DETERMINISTIC scatter (entries of a
point added in ascending (n, k) order, csrc/train_ops.cu) for float32; MVPNET_B200_DETERMINISTIC=0 (or float64) selects
the atomicAdd kernel, whose summation order — like the reference's — changes from run to run."""
import torch

from ._util import deterministic, ext


class GroupPointsFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, index):
        ctx.save_for_backward(index)
        ctx.num_points = points.size(2)
        return ext().group_points_cuda.group_points_forward(points, index)

    @staticmethod
    def backward(ctx, *grad_output):
        (index,) = ctx.saved_tensors
        g = grad_output[0]
        fn = ext().group_points_cuda.group_points_backward_det if deterministic(g) else ext().group_points_cuda.group_points_backward
        grad = fn(g, index, ctx.num_points)
        return grad, None


def group_points(points, index):
    """points (B, C, N1), index int64 (B, N2, K) -> (B, C, N2, K) with out[b,c,n,k] = points[b,c,index[b,n,k]]."""
    return GroupPointsFunction.apply(points, index)
