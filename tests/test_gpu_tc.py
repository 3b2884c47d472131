"""GPU: the tcgen05 (bf16 hi/lo x 3 products) fused kernels against the fp32 SIMT fused kernels and the
golden logits, per stage and end to end.  Tolerance 1e-4 of max|reference| (north_star)."""
import os

import numpy as np
import pytest
import torch

from mvpnet_b200 import synthetic

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


@pytest.fixture(autouse=True)
def strict():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        yield


def run_backend(backend, fn):
    from mvpnet_b200 import engine
    old = engine.MLP_BACKEND
    engine.MLP_BACKEND = backend
    try:
        return fn()
    finally:
        engine.MLP_BACKEND = old


@pytest.mark.parametrize('b', [1, 3])
def test_pn2_full_tc_vs_simt_vs_golden(b):
    from mvpnet_b200 import engine
    from mvpnet_b200.modules import PN2SSG
    pts, _ = synthetic.room_points(8192, seed=0)
    feat = torch.randn(1, 64, 8192, generator=torch.Generator().manual_seed(13))
    net = synthetic.fill_parameters(PN2SSG(64, 20), seed=5).eval().cuda()
    xyz = torch.from_numpy(pts.T.copy())[None].cuda().repeat(b, 1, 1)
    batch = {'points': xyz, 'feature': feat.cuda().repeat(b, 1, 1)}
    tc = run_backend('tc', lambda: net.fast_forward(batch)['seg_logit'])
    sa_chains, fp_chains = engine._CACHE[(id(net), 'pn2tc')][1]
    assert all(isinstance(c, engine.TcChain) for c in sa_chains), 'tensor-core path was not selected'
    simt = run_backend('simt', lambda: net.fast_forward(batch)['seg_logit'])
    g = np.load(os.path.join(GOLD, 'pn2_full.npz'))['logit']
    assert rel(simt[0].cpu(), torch.from_numpy(g[0])) < 1e-4
    assert rel(tc[0].cpu(), torch.from_numpy(g[0])) < 1e-4
    assert rel(tc, simt) < 1e-4
    if b > 1:
        assert torch.equal(tc[0], tc[b - 1])      # identical clouds in the batch give identical rows


def test_stagewise_tc_vs_simt():
    """Each fused stage separately, random point-major inputs, including ragged tile counts."""
    import mvpnet_b200
    from mvpnet_b200 import engine
    from mvpnet_b200.modules import SharedMLP
    ext = mvpnet_b200.load_ext()
    torch.manual_seed(3)
    dev = 'cuda'
    # set abstraction: B=2, N=1000, M=37 (ragged: 37 centroids -> 10 tiles, last one partial), C=64
    B, N, M, C = 2, 1000, 37, 64
    mlp = synthetic.fill_parameters(SharedMLP(C + 3, (64, 96, 128), ndim=2), seed=1).eval().to(dev)
    feat = torch.randn(B, N, C, device=dev)
    xyz = torch.rand(B, N, 3, device=dev)
    new_xyz = xyz[:, :M].contiguous()
    nbr = torch.randint(0, N, (B, M, 32), device=dev)
    nbr[0, 3, 5:] = -1                                   # out-of-range rows are zero rows in both kernels
    simt = engine.Chain(engine._mlp_layers(mlp), C + 3, dev)
    tc = engine.TcChain(engine._mlp_layers(mlp), C + 3, dev)
    a = ext.fused_cuda.set_abstraction(feat, xyz, new_xyz, nbr, *simt.args())
    b = ext.fused_cuda.tc_set_abstraction(feat, xyz, new_xyz, nbr, *tc.args())
    assert rel(b, a) < 2e-5
    # xyz-only set abstraction (3D-only network: in_channels = 0)
    mlp0 = synthetic.fill_parameters(SharedMLP(3, (32, 32, 64), ndim=2), seed=2).eval().to(dev)
    a = ext.fused_cuda.set_abstraction(None, xyz, new_xyz, nbr, *engine.Chain(engine._mlp_layers(mlp0), 3, dev).args())
    b = ext.fused_cuda.tc_set_abstraction(None, xyz, new_xyz, nbr, *engine.TcChain(engine._mlp_layers(mlp0), 3, dev).args())
    assert rel(b, a) < 2e-5
    # feature propagation: Ns=50, Nd=333 (3 tiles, ragged), Cs=128, Cd=64, 5-layer chain with a linear tail
    Ns, Nd, Cs, Cd = 50, 333, 128, 64
    mlp1 = synthetic.fill_parameters(SharedMLP(Cs + Cd, (128, 64), ndim=1), seed=4).eval().to(dev)
    head = synthetic.fill_parameters(torch.nn.Conv1d(64, 20, 1), seed=5).to(dev)
    layers = engine._mlp_layers(mlp1) + [(head, None, False)]
    sparse = torch.randn(B, Ns, Cs, device=dev)
    skip = torch.randn(B, Nd, Cd, device=dev)
    idx = torch.randint(0, Ns, (B, Nd, 3), device=dev)
    d2 = torch.rand(B, Nd, 3, device=dev) + 1e-3
    a = ext.fused_cuda.feature_propagation(sparse, idx, d2, skip, 1e-10, *engine.Chain(layers, Cs + Cd, dev).args())
    b = ext.fused_cuda.tc_feature_propagation(sparse, idx, d2, skip, 1e-10, *engine.TcChain(layers, Cs + Cd, dev).args())
    assert tuple(b.shape) == (B, Nd, 20) and rel(b, a) < 2e-5
    # feature aggregation: channels-last and NCHW feature maps, k = 3 sum and k = 2 max
    nv, h, w, Np = 2, 12, 16, 77
    fa_mlp = synthetic.fill_parameters(SharedMLP(68, (64, 64, 64), ndim=2), seed=6).eval().to(dev)
    f2d = torch.randn(B, nv, 64, h, w, device=dev)
    f2d_cl = f2d.reshape(B * nv, 64, h, w).contiguous(memory_format=torch.channels_last).view(B, nv, 64, h, w)
    pix = torch.rand(B, nv * h * w, 3, device=dev)
    pts = torch.rand(B, Np, 3, device=dev)
    for k, red in ((3, True), (2, False)):
        knn = torch.randint(0, nv * h * w, (B, Np, k), device=dev)
        a = ext.fused_cuda.feature_aggregation(f2d, pix, pts, knn, red, *engine.Chain(engine._mlp_layers(fa_mlp), 68, dev).args())
        for fm in (f2d, f2d_cl):
            b = ext.fused_cuda.tc_feature_aggregation(fm, pix, pts, knn, red, *engine.TcChain(engine._mlp_layers(fa_mlp), 68, dev).args())
            assert rel(b, a) < 2e-5
