import torch

from .. import load_ext


def ext():
    return load_ext()


def channels_last(x, transpose):
    """(B, 3, N) -> contiguous (B, N, 3) when `transpose`, else just contiguous."""
    if transpose:
        x = x.transpose(1, 2)
    return x.contiguous()


class _NoGrad(torch.autograd.Function):
    """Base for the index-producing ops: outputs are integer/auxiliary, every input grad is None
    (reference: fps.py:11-13, ball_query.py:12-14, knn_distance.py:11-13)."""

    @staticmethod
    def backward(ctx, *grad_outputs):
        return (None,) * ctx.num_inputs
