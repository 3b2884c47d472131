"""GPU parity against golden vectors produced by the REFERENCE's own Python (tests/golden/
make_golden.py: unmodified /root/reference modules on CPU, extension ops served by the oracle).
Inputs are regenerated from seeds (mvpnet_b200.synthetic) and checked against stored checksums.

Tolerances (north_star): indices bit-exact; features / logits within 1e-4 relative.  Two figures are checked
and printed for every float comparison:
  tensor-normalised   max|a - b| / max|b|                         < 1e-4
  element-wise        max_i |a_i - b_i| / (|b_i| + rms(b))        < 1e-4   (rms(b) is the absolute floor: logits
                      cross zero, so a pure |a_i - b_i| / |b_i| is unbounded for any finite-precision implementation)"""
import os

import numpy as np
import pytest
import torch

from mvpnet_b200 import synthetic

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
REL_TOL = 1e-4

PN2_SMALL = dict(sa_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128, 256)),
                 num_centroids=(512, 128, 32, 8), radius=(0.2, 0.4, 0.8, 1.6), max_neighbors=(32, 32, 32, 32),
                 fp_channels=((128, 128), (128, 128), (128, 64), (64, 64, 64)), seg_channels=(64,))


@pytest.fixture(autouse=True)
def strict_fp32():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def checksum(a):
    return float(np.asarray(a, np.float64).sum())


def rel_err(got, want):
    """max of the tensor-normalised and the element-wise (rms-floored) relative error; prints both."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    diff = np.abs(got - want)
    tensor = diff.max() / np.abs(want).max()
    elem = (diff / (np.abs(want) + np.sqrt((want ** 2).mean()))).max()
    print('rel err: tensor-normalised %.2e, element-wise (rms floor) %.2e' % (tensor, elem))
    return max(tensor, elem)


def rgbd_on_gpu(chunk, k=3):
    """unproject + 2D->3D k-NN through the public data-side API."""
    from mvpnet_b200.data import unproject_and_knn
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return unproject_and_knn(t(chunk['depth'])[None], t(chunk['cam_matrix'])[None], t(chunk['pose'])[None],
                             t(chunk['points'])[None], k=k, chunk_box=t(chunk['chunk_box'])[None])


def test_rgbd_chunk_matches_reference_get_rgbd_data():
    g = np.load(os.path.join(GOLD, 'rgbd_chunk.npz'))
    chunk = synthetic.make_chunk(seed=0)
    assert checksum(chunk['depth']) + checksum(chunk['pose']) + checksum(chunk['points']) == float(g['in_checksum'])
    out = rgbd_on_gpu(chunk)
    want_mask = np.unpackbits(g['image_mask'])[:5 * 120 * 160].reshape(5, 120, 160).astype(bool)
    assert np.array_equal(out['image_mask'][0].cpu().numpy().astype(bool), want_mask)
    assert np.array_equal(out['image_xyz'][0].cpu().numpy(), g['image_xyz'])
    assert np.array_equal(out['knn_indices'][0].cpu().numpy(), g['knn_indices'].astype(np.int64))


def test_fa_c1_config1():
    from mvpnet_b200.modules import FeatureAggregation
    from mvpnet_b200.ops import group_points
    g = np.load(os.path.join(GOLD, 'fa_c1.npz'))
    c1 = synthetic.make_chunk(seed=1, num_points=2048, num_views=1)
    feat2d = torch.randn(1, 64, 1, 120, 160, generator=torch.Generator().manual_seed(11))
    assert checksum(feat2d.numpy()) + checksum(c1['points']) == float(g['in_checksum'])
    out = rgbd_on_gpu(c1)
    assert np.array_equal(out['knn_indices'][0].cpu().numpy(), g['knn_indices'].astype(np.int64))
    fa = synthetic.fill_parameters(FeatureAggregation(64), seed=3).eval().cuda()
    knn = out['knn_indices']
    f = group_points(feat2d.cuda().reshape(1, 64, -1), knn)
    xyz = group_points(out['image_xyz'].permute(0, 4, 1, 2, 3).reshape(1, 3, -1), knn)
    pts = torch.from_numpy(c1['points'].T.copy())[None].cuda()
    y = fa(xyz, pts, f)
    assert rel_err(y.cpu().numpy(), g['out']) < REL_TOL
    from mvpnet_b200 import engine
    y2 = engine.feature_aggregation(fa, feat2d.cuda().reshape(1, 1, 64, 120, 160), out['image_xyz'], knn, pts)
    assert rel_err(y2.cpu().numpy(), g['out']) < REL_TOL


@pytest.mark.parametrize('fast', [False, True])
def test_pn2_small(fast):
    from mvpnet_b200.modules import PN2SSG
    from mvpnet_b200.ops import ball_query, farthest_point_sample
    g = np.load(os.path.join(GOLD, 'pn2_small.npz'))
    pts, _ = synthetic.room_points(2048, seed=2)
    feat = torch.randn(1, 16, 2048, generator=torch.Generator().manual_seed(12))
    assert checksum(pts) + checksum(feat.numpy()) == float(g['in_checksum'])
    xyz = torch.from_numpy(pts.T.copy())[None].cuda()
    cur = xyz
    for i, (m, r) in enumerate(zip(PN2_SMALL['num_centroids'], PN2_SMALL['radius'])):
        idx = farthest_point_sample(cur, m)
        assert np.array_equal(idx[0].cpu().numpy(), g['fps%d' % i].astype(np.int64)), 'fps level %d' % i
        new = torch.gather(cur, 2, idx[:, None, :].expand(1, 3, m))
        bq = ball_query(new, cur, r, 32)
        assert np.array_equal(bq[0].cpu().numpy(), g['bq%d' % i].astype(np.int64)), 'ball_query level %d' % i
        cur = new
    net = synthetic.fill_parameters(PN2SSG(16, 20, **PN2_SMALL), seed=4).eval().cuda()
    batch = {'points': xyz, 'feature': feat.cuda()}
    logit = net.fast_forward(batch)['seg_logit'] if fast else net(batch)['seg_logit']
    assert rel_err(logit.cpu().numpy(), g['logit']) < REL_TOL


@pytest.mark.parametrize('fast', [False, True])
def test_pn2_full_config2_forward(fast):
    from mvpnet_b200.modules import PN2SSG
    g = np.load(os.path.join(GOLD, 'pn2_full.npz'))
    pts, _ = synthetic.room_points(8192, seed=0)
    feat = torch.randn(1, 64, 8192, generator=torch.Generator().manual_seed(13))
    assert checksum(pts) + checksum(feat.numpy()) == float(g['in_checksum'])
    net = synthetic.fill_parameters(PN2SSG(64, 20), seed=5).eval().cuda()
    batch = {'points': torch.from_numpy(pts.T.copy())[None].cuda(), 'feature': feat.cuda()}
    logit = net.fast_forward(batch)['seg_logit'] if fast else net(batch)['seg_logit']
    assert rel_err(logit.cpu().numpy(), g['logit']) < REL_TOL


@pytest.mark.parametrize('fast', [False, True, 'from_depth'])
def test_mvpnet_config3(fast):
    from mvpnet_b200.modules import MVPNet3D, PN2SSG
    from mvpnet_b200.unet import UNetResNet34
    g = np.load(os.path.join(GOLD, 'mvpnet_c3.npz'))
    chunk = synthetic.make_chunk(seed=0)
    model = MVPNet3D(UNetResNet34(20, p=0.5, pretrained=False), None, PN2SSG(64, 20), in_channels=64,
                     mlp_channels=(64, 64, 64), reduction='sum', use_relation=True)
    synthetic.fill_parameters(model, seed=6).eval().cuda()
    images = torch.from_numpy(chunk['images'])[None].cuda()
    feat2d = model.net_2d({'image': images[0]})['feature']
    assert abs(checksum(feat2d.cpu().numpy()) - float(g['feat2d_checksum'])) < 1e-4 * abs(float(g['feat2d_checksum'])) + 50
    assert rel_err(feat2d[:, :, ::16, ::16].cpu().numpy(), g['feat2d_sample']) < REL_TOL
    rg = rgbd_on_gpu(chunk)
    batch = {'images': images, 'image_xyz': rg['image_xyz'], 'knn_indices': rg['knn_indices'],
             'points': torch.from_numpy(chunk['points'].T.copy())[None].cuda()}
    if fast == 'from_depth':      # data side (unproject + 3-NN) inside the fused forward, on the side stream
        from mvpnet_b200.data import invert_intrinsics
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        raw = {'images': images, 'points': batch['points'], 'depth': t(chunk['depth'])[None], 'pose': t(chunk['pose'])[None],
               'cam_inv': t(np.broadcast_to(invert_intrinsics(chunk['cam_matrix']), (5, 3, 3)).copy())[None],
               'chunk_box': t(chunk['chunk_box'])[None], 'k': 3}
        logit = model.fast_forward(raw)['seg_logit']
    else:
        logit = model.fast_forward(batch)['seg_logit'] if fast else model(batch)['seg_logit']
    assert rel_err(logit.cpu().numpy(), g['logit']) < REL_TOL
