"""GPU: this package's kernels against the REFERENCE'S OWN CUDA kernels, compiled unmodified for
sm_100a from /root/reference/mvpnet/ops/cuda by oracle/build_ref.py into oracle/_ref/ (the only
injected piece is a <THC/THC.h> compatibility header).  This pins both the sm_100a kernels and the
CPU oracle to the real reference on the reference's own test shapes and on the model's shapes.
Skipped when oracle/_ref/ was not built (it needs /root/reference at build time)."""
import glob
import importlib.util
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref')


def load_ref(name):
    hits = glob.glob(os.path.join(REF_DIR, name + '*.so'))
    if not hits:
        pytest.skip('oracle/_ref/%s not built' % name)
    spec = importlib.util.spec_from_file_location(name, hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope='module')
def ext():
    import mvpnet_b200
    return mvpnet_b200.load_ext()


def room(b, n, seed):
    from mvpnet_b200 import synthetic
    return np.stack([synthetic.room_points(n, seed + i)[0] for i in range(b)])


@pytest.mark.parametrize('b,d,n,m,dtype', [(2, 3, 1024, 128, torch.float64), (2, 2, 1024, 128, torch.float64),
                                           (3, 3, 1025, 129, torch.float32), (4, 3, 8192, 2048, torch.float32),
                                           (2, 3, 2048, 512, torch.float32), (2, 3, 128, 32, torch.float32)])
def test_fps_vs_reference_kernel(ext, b, d, n, m, dtype):
    ref = load_ref('fps_cuda')
    np.random.seed(0)
    pts = torch.from_numpy(np.random.rand(b, n, d)).to(dtype).cuda()
    want = ref.farthest_point_sample(pts, m)
    got = ext.fps_cuda.farthest_point_sample(pts, m)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    assert np.array_equal(oracle.farthest_point_sample(pts.cpu().numpy(), m), want.cpu().numpy())


def test_fps_ties_vs_reference_kernel(ext):
    """Duplicated / lattice points: the tie rule of the reference launch geometry, on the real kernel."""
    ref = load_ref('fps_cuda')
    pts = room(3, 8192, 11)
    pts[1, 4096:] = pts[1, :4096]
    pts[2] = np.round(pts[2] * 20) / 20
    t = torch.from_numpy(pts).cuda()
    want = ref.farthest_point_sample(t, 2048)
    got = ext.fps_cuda.farthest_point_sample(t, 2048)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    assert np.array_equal(oracle.farthest_point_sample(pts, 2048), want.cpu().numpy())
    small = np.round(room(2, 128, 3) * 4) / 4          # BLOCK = 128 path, heavy ties
    t = torch.from_numpy(small).cuda()
    assert torch.equal(ext.fps_cuda.farthest_point_sample(t, 32), ref.farthest_point_sample(t, 32))


def tie_cloud(n, q, seed):
    """Lattice cloud with heavy exact ties; second half duplicates the first (CropPad-style)."""
    rng = np.random.RandomState(seed)
    p = np.round(rng.rand(n, 3).astype(np.float32) * np.array([1.9, 1.9, 2.5], np.float32) * q) / q
    p[n // 2:n // 2 * 2] = p[:n // 2]
    return p.astype(np.float32)


@pytest.mark.parametrize('n', [300, 1000, 1024, 2048, 3000, 4096, 5000, 8192])
@pytest.mark.parametrize('q', [4, 8, 20])
def test_fps_tie_order_all_geometries(ext, n, q):
    """the reference tie order (bit-reversed thread of a BLOCK = min(2^floor(log2 N), 512)
    launch, fps_kernel.cu:95-129) at every launch geometry of this package's kernel, M up to N/2, on the
    reference's own kernel and on the oracle."""
    ref = load_ref('fps_cuda')
    pts = np.stack([tie_cloud(n, q, 100 * q + i) for i in range(2)])
    t = torch.from_numpy(pts).cuda()
    m = n // 2
    want = ref.farthest_point_sample(t, m)
    got = ext.fps_cuda.farthest_point_sample(t, m)
    torch.cuda.synchronize()
    assert torch.equal(got, want), 'first mismatch at step %d' % int((got != want).any(0).nonzero()[0])
    assert np.array_equal(oracle.farthest_point_sample(pts, m), want.cpu().numpy())


@pytest.mark.parametrize('b,n1,n2,r,k,dtype', [(2, 64, 128, 0.1, 32, torch.float64), (3, 65, 129, 10.0, 32, torch.float64),
                                               (3, 65, 129, 0.1, 32, torch.float32), (4, 512, 1024, 0.1, 64, torch.float32)])
def test_ball_query_vs_reference_kernel(ext, b, n1, n2, r, k, dtype):
    ref, refd = load_ref('ball_query_cuda'), load_ref('ball_query_distance_cuda')
    np.random.seed(0)
    key = np.random.randn(b, n2, 3)
    query = np.stack([p[np.random.choice(n2, n1, replace=False)] for p in key])
    q, kk = torch.from_numpy(query).to(dtype).cuda(), torch.from_numpy(key).to(dtype).cuda()
    want = ref.ball_query(q, kk, r, k)
    wi, wd = refd.ball_query_distance(q, kk, r, k)
    got = ext.ball_query_cuda.ball_query(q, kk, r, k)
    gi, gd = ext.ball_query_distance_cuda.ball_query_distance(q, kk, r, k)
    torch.cuda.synchronize()
    assert torch.equal(got, want) and torch.equal(gi, wi) and torch.equal(gd, wd)
    assert np.array_equal(oracle.ball_query(q.cpu().numpy(), kk.cpu().numpy(), r, k), want.cpu().numpy())


def test_model_levels_vs_reference_kernels(ext):
    """ball_query / knn_distance / group_points / interpolate at the four PN2SSG levels, B = 4."""
    rbq, rknn = load_ref('ball_query_cuda'), load_ref('knn_distance_cuda')
    rgp, rip = load_ref('group_points_cuda'), load_ref('interpolate_cuda')
    rfps = load_ref('fps_cuda')
    cur = torch.from_numpy(room(4, 8192, 21)).cuda()
    torch.manual_seed(0)
    levels = [cur]
    for m, r in [(2048, 0.1), (512, 0.2), (128, 0.4), (32, 0.8)]:
        idx = rfps.farthest_point_sample(cur, m)
        assert torch.equal(ext.fps_cuda.farthest_point_sample(cur, m), idx)
        new = torch.gather(cur, 1, idx.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        want = rbq.ball_query(new, cur, r, 32)
        got = ext.ball_query_cuda.ball_query(new, cur, r, 32)
        assert torch.equal(got, want)
        feat = torch.randn(4, 16, cur.size(1), device='cuda')
        assert torch.equal(ext.group_points_cuda.group_points_forward(feat, got), rgp.group_points_forward(feat, want))
        go = torch.randn(4, 16, m, 32, device='cuda')
        a = ext.group_points_cuda.group_points_backward(go, got, cur.size(1))
        bref = rgp.group_points_backward(go, want, cur.size(1))
        assert torch.allclose(a, bref, rtol=1e-4, atol=1e-4)
        levels.append(new)
        cur = new
    for i in range(4):
        dense, sparse = levels[3 - i], levels[4 - i]
        wi, wd = rknn.knn_distance(dense, sparse, 3)
        gi, gd = ext.knn_distance_cuda.knn_distance(dense, sparse, 3)
        assert torch.equal(gi, wi) and torch.equal(gd, wd)
        feat = torch.randn(4, 32, sparse.size(1), device='cuda')
        w = torch.rand(4, dense.size(1), 3, device='cuda')
        assert torch.equal(ext.interpolate_cuda.interpolate_forward(feat, gi, w), rip.interpolate_forward(feat, wi, w))
        go = torch.randn(4, 32, dense.size(1), device='cuda')
        assert torch.allclose(ext.interpolate_cuda.interpolate_backward(go, gi, w, sparse.size(1)),
                              rip.interpolate_backward(go, wi, w, sparse.size(1)), rtol=1e-4, atol=1e-4)
    torch.cuda.synchronize()


@pytest.mark.parametrize('b,n1,n2', [(2, 512, 1024), (3, 513, 1025), (3, 31, 63)])
def test_knn_vs_reference_kernel(ext, b, n1, n2):
    ref = load_ref('knn_distance_cuda')
    torch.manual_seed(0)
    q, k = torch.randn(b, n1, 3).cuda(), torch.randn(b, n2, 3).cuda()
    wi, wd = ref.knn_distance(q, k, 3)
    gi, gd = ext.knn_distance_cuda.knn_distance(q, k, 3)
    torch.cuda.synchronize()
    assert torch.equal(gi, wi) and torch.equal(gd, wd)
    oi, od = oracle.knn_distance(q.cpu().numpy(), k.cpu().numpy(), 3)
    assert np.array_equal(oi, wi.cpu().numpy()) and np.array_equal(od, wd.cpu().numpy())
