"""Scene-level callers (mvpnet_b200/scene.py) against the golden fixture produced by the reference's own
chunk_util.scene2chunks_legacy and the vote accumulation of test_mvpnet_3d.py (tests/golden/make_golden_scene.py).
Index lists, counts and labels must be identical; mean logits bit-identical (same fp32 adds in the same order)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from mvpnet_b200 import scene

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location('make_golden_scene_inputs', os.path.join(HERE, 'golden', 'make_golden_scene.py'))


def _scene_points():
    # the generator's point sampler without importing the reference (absent on the GPU box)
    src = open(os.path.join(HERE, 'golden', 'make_golden_scene.py')).read()
    ns = {}
    start, end = src.index('def scene_points'), src.index('def reference_select_frames')
    from mvpnet_b200 import synthetic
    exec(src[start:end], {'np': np, 'synthetic': synthetic}, ns)
    return ns['scene_points']()


def _overlap_matrix():
    src = open(os.path.join(HERE, 'golden', 'make_golden_scene.py')).read()
    ns = {}
    start, end = src.index('def overlap_matrix'), src.index('def main')
    exec(src[start:end], {'np': np}, ns)
    return ns['overlap_matrix']()


def _check_frames(device):
    g = np.load(os.path.join(HERE, 'golden', 'select_frames.npz'))
    m = torch.from_numpy(_overlap_matrix()).to(device)
    keep = m.clone()
    for k in (1, 3, 5, 12):
        assert scene.select_frames(m, k) == g[str(k)].tolist()
    assert torch.equal(m, keep)                   # input untouched


def test_select_frames_cpu():
    _check_frames('cpu')


@pytest.mark.gpu
def test_select_frames_gpu():
    _check_frames('cuda')


def _check(device):
    g = np.load(os.path.join(HERE, 'golden', 'scene_chunks.npz'))
    pts = _scene_points()
    assert abs(float(pts.astype(np.float64).sum()) - float(g['points_checksum'])) < 1e-6
    p = torch.from_numpy(pts).to(device)
    idx, bbox = scene.scene2chunks_legacy(p, chunk_size=(1.5, 1.5), stride=0.5, thresh=1000, margin=(0.2, 0.2), return_bbox=True)
    assert len(idx) == int(g['num_chunks'])
    assert [int(i.numel()) for i in idx] == g['chunk_sizes'].tolist()
    assert [int(i.sum()) for i in idx] == g['chunk_index_checksums'].tolist()
    assert np.array_equal(idx[0].cpu().numpy(), g['first_chunk'])
    # corners: float64 here (= the reference under its NumPy 1.x, where float32 scalar + Python float promotes to
    # float64); the fixture was produced under NumPy 2.3, whose NEP 50 keeps float32 corners: equal to ~1e-7
    np.testing.assert_allclose(torch.stack(bbox).cpu().numpy(), g['bboxes'], rtol=0, atol=1e-6)
    acc = scene.VoteAccumulator(len(pts), 20, device)
    for c, ind in enumerate(idx):
        logit = torch.from_numpy(np.random.RandomState(1000 + c).randn(20, ind.numel() + 7).astype(np.float32)).to(device)
        acc.add(ind, logit)
    mean, label = acc.finalize()
    assert np.array_equal(acc.count.cpu().numpy(), g['count'].astype(np.int32))
    assert (g['count'] == 0).sum() > 0 and np.array_equal(label.cpu().numpy(), g['label'])
    assert np.array_equal(mean[::97].cpu().numpy(), g['mean_sample'])


def test_scene_chunks_and_votes_cpu():
    _check('cpu')


@pytest.mark.gpu
def test_scene_chunks_and_votes_gpu():
    _check('cuda')


@pytest.mark.gpu
def test_mvpnet2d_fast_matches_module():
    import warnings
    from mvpnet_b200 import synthetic
    from mvpnet_b200.modules import MVPNet2D
    from mvpnet_b200.unet import UNetResNet34
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        model = MVPNet2D(UNetResNet34(20, p=0.5, pretrained=False))
    synthetic.fill_parameters(model, seed=8).eval().cuda()
    b, nv, h, w, npts = 2, 3, 120, 160, 4096
    g = torch.Generator().manual_seed(1)
    batch = {'images': torch.randn(b, nv, 3, h, w, generator=g).cuda(),
             'knn_indices': torch.randint(0, nv * h * w, (b, npts, 3), generator=g).cuda()}
    with torch.no_grad():
        want = model(batch)['seg_logit']
        got = model.fast_forward(batch)['seg_logit']
    assert got.shape == want.shape
    assert float((got - want).abs().max() / want.abs().max()) < 1e-4
