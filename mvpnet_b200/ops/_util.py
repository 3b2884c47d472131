import os

import torch

from .. import load_ext


def ext():
    return load_ext()


DETERMINISTIC = os.environ.get('MVPNET_B200_DETERMINISTIC', '1') == '1'


def deterministic(grad):
    """Use the fixed-order backward scatter (float32 only) instead of atomicAdd."""
    return DETERMINISTIC and grad.dtype == torch.float32


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
