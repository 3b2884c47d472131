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
        picks = scene.select_frames_device(m, k)                 # no host round trip per pick
        assert picks.device == m.device and picks.tolist() == g[str(k)].tolist()
    assert torch.equal(m, keep)                   # input untouched


def _golden_fn(name, until):
    src = open(os.path.join(HERE, 'golden', 'make_golden_scene.py')).read()
    ns = {}
    exec(src[src.index('def ' + name):src.index(until)], {'np': np}, ns)
    return ns[name]


def _check_boundary(device):
    """Extent exactly on a window boundary + points exactly on window edges (limit must be subtracted in the
    points' dtype and the corners follow the NumPy promotion the fixture was generated under)."""
    g = np.load(os.path.join(HERE, 'golden', 'scene_boundary.npz'))
    pts = _golden_fn('boundary_points', 'def mvpnet2d_golden')()
    assert abs(float(pts.astype(np.float64).sum()) - float(g['points_checksum'])) < 1e-6
    promotion = 'nep50' if int(str(g['numpy_version']).split('.')[0]) >= 2 else 'legacy'
    idx = scene.scene2chunks_legacy(torch.from_numpy(pts).to(device), chunk_size=(1.5, 1.5), stride=0.5, thresh=200, margin=(0.2, 0.2),
                                    promotion=promotion)
    assert [int(i.numel()) for i in idx] == g['chunk_sizes'].tolist()
    assert [int(i.sum()) for i in idx] == g['chunk_index_checksums'].tolist()


def test_scene_chunks_boundary_cpu():
    _check_boundary('cpu')


@pytest.mark.gpu
def test_scene_chunks_boundary_gpu():
    _check_boundary('cuda')


@pytest.mark.gpu
def test_nearest_propagation_matches_sklearn():
    """1-NN label propagation (test_3d_scene.py:155-163) with this package's k = 1 grid search against scikit-learn."""
    g = np.load(os.path.join(HERE, 'golden', 'nearest_1nn.npz'))
    pts = torch.from_numpy(_scene_points()[:40000]).cuda()
    ind = torch.from_numpy(g['vote_indices'].astype(np.int64)).cuda()
    vp = pts.index_select(0, ind.reshape(-1)).reshape(2, 8192, 3)
    # logits that encode the sample id: the propagated "logit" of a scene point reveals which sample it took
    ids = torch.arange(8192, dtype=torch.float32, device='cuda')
    logits = torch.stack([torch.stack([ids, -ids]), torch.stack([ids * 2, ids])])            # (v=2, c=2, m)
    mean, label = scene.propagate_nearest(pts, vp, logits)
    nn = torch.from_numpy(g['nn_indices'].astype(np.int64)).cuda()
    want = (nn[0].float() + 2 * nn[1].float()) / 2
    assert torch.equal(mean[:, 0], want)
    assert mean.shape == (40000, 2) and label.shape == (40000,)


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
    # corners follow the NumPy >= 2 promotion the fixture was produced under (float32 sums): bit-equal
    assert np.array_equal(torch.stack(bbox).cpu().numpy(), g['bboxes'].astype(np.float64))
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


@pytest.mark.gpu
def test_mvpnet2d_matches_reference_golden():
    """MVPNet2D (mvpnet/models/mvpnet_2d.py:7-34) module and fast path against the reference's own Python
    (tests/golden/make_golden_scene.py::mvpnet2d_golden), 1e-4 of max."""
    import warnings
    from mvpnet_b200 import synthetic
    from mvpnet_b200.modules import MVPNet2D
    from mvpnet_b200.unet import UNetResNet34
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(os.path.join(HERE, 'golden', 'mvpnet2d.npz'))
    knn = np.load(os.path.join(HERE, 'golden', 'rgbd_chunk.npz'))['knn_indices'].astype(np.int64)
    chunk = synthetic.make_chunk(seed=0)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        model = MVPNet2D(UNetResNet34(20, p=0.5, pretrained=False))
    synthetic.fill_parameters(model, seed=8).eval().cuda()
    batch = {'images': torch.from_numpy(chunk['images'])[None].cuda(), 'knn_indices': torch.from_numpy(knn)[None].cuda()}
    with torch.no_grad():
        for out in (model(batch)['seg_logit'], model.fast_forward(batch)['seg_logit']):
            got = out[0, :, ::4].cpu().numpy().astype(np.float64)
            assert np.abs(got - g['logit_sample']).max() / float(g['logit_absmax']) < 1e-4
