"""GPU: BASELINE config 2 backward — PN2SSG in TRAIN mode (batch-statistics BatchNorm) through the op-by-op
modules (sm_100a forward ops + the atomic scatter backward of group_points / feature_interpolate), against
gradients produced by the reference's own PN2SSG on CPU (tests/golden/make_golden.py).  Tolerance 1e-3 of the
tensor's max: the reference's own backward is order-nondeterministic (atomicAdd) and BN statistics are reduced in
a different order on cuDNN."""
import os

import numpy as np
import pytest
import torch

from mvpnet_b200 import synthetic

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
PN2_SMALL = dict(sa_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128, 256)),
                 num_centroids=(512, 128, 32, 8), radius=(0.2, 0.4, 0.8, 1.6), max_neighbors=(32, 32, 32, 32),
                 fp_channels=((128, 128), (128, 128), (128, 64), (64, 64, 64)), seg_channels=(64,))


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / np.abs(b).max()


def test_pn2_train_forward_backward_matches_reference():
    from mvpnet_b200.modules import PN2SSG
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(os.path.join(GOLD, 'pn2_small_train.npz'))
    pts, _ = synthetic.room_points(2048, seed=2)
    feat = torch.randn(1, 16, 2048, generator=torch.Generator().manual_seed(12)).cuda().requires_grad_(True)
    net = synthetic.fill_parameters(PN2SSG(16, 20, dropout_prob=0.0, **PN2_SMALL), seed=4).train().cuda()
    xyz = torch.from_numpy(pts.T.copy())[None].cuda()
    logit = net({'points': xyz, 'feature': feat})['seg_logit']
    assert rel(logit.detach().cpu().numpy(), g['logit']) < 1e-3
    wgt = torch.randn(logit.shape, generator=torch.Generator().manual_seed(14)).cuda()
    (logit * wgt).sum().backward()
    assert rel(feat.grad.cpu().numpy(), g['feat_grad']) < 1e-3
    params = dict(net.named_parameters())
    for name in ('sa_modules.0.mlp.0.conv.weight', 'sa_modules.3.mlp.2.conv.weight', 'fp_modules.0.mlp.0.conv.weight',
                 'fp_modules.3.mlp.2.bn.weight', 'seg_logit.weight'):
        assert rel(params[name].grad.cpu().numpy(), g['g_' + name.replace('.', '_')]) < 1e-3, name
    with pytest.raises(RuntimeError, match='inference-only'):
        net.fast_forward({'points': xyz, 'feature': feat.detach()})
