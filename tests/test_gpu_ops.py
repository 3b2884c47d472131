"""GPU parity: the sm_100a extension (through the torch shim over the C ABI) against the CPU oracle
on the same seeded inputs.  Index outputs must be bit-exact; float outputs bit-exact where the
arithmetic is fully specified (distances, forward interpolate), tolerance where the reference
itself is order-nondeterministic (atomic backward)."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ext():
    import mvpnet_b200
    return mvpnet_b200.load_ext()


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def room_points(b, n, seed=0):
    """Synthetic room: points on floor / walls / boxes of a 1.9 x 1.9 x 2.5 m chunk."""
    rng = np.random.RandomState(seed)
    out = np.empty((b, n, 3), np.float32)
    for i in range(b):
        u = rng.rand(n, 3).astype(np.float32) * np.array([1.9, 1.9, 2.5], np.float32)
        sel = rng.rand(n)
        u[sel < 0.35, 2] = 0.0
        w1 = (sel >= 0.35) & (sel < 0.55)
        u[w1, 0] = 0.0
        w2 = (sel >= 0.55) & (sel < 0.75)
        u[w2, 1] = 1.9
        box = sel >= 0.75
        u[box] = u[box] * 0.3 + np.array([0.6, 0.6, 0.0], np.float32)
        out[i] = u + rng.randn(n, 3).astype(np.float32) * 0.005
    return out


# ---------------------------------------------------------------- FPS
@pytest.mark.parametrize('b,d,n,m,dtype', [
    (2, 3, 1024, 128, np.float64), (2, 2, 1024, 128, np.float64), (3, 3, 1025, 129, np.float64),
    (2, 3, 1024, 128, np.float32), (2, 2, 1000, 100, np.float32), (3, 3, 1025, 129, np.float32),
    (4, 3, 8192, 2048, np.float32), (2, 3, 2048, 512, np.float32), (2, 3, 512, 128, np.float32),
    (2, 3, 128, 32, np.float32), (1, 3, 37, 37, np.float32), (1, 3, 5, 3, np.float32),
    (1, 3, 9000, 300, np.float32), (1, 3, 20000, 64, np.float64)])
def test_fps(ext, b, d, n, m, dtype):
    np.random.seed(0)
    pts = np.random.rand(b, n, d).astype(dtype)
    want = oracle.farthest_point_sample(pts, m)
    got = ext.fps_cuda.farthest_point_sample(cu(pts), m)
    assert got.dtype == torch.int64 and tuple(got.shape) == (b, m)
    assert np.array_equal(got.cpu().numpy(), want)


def test_fps_room_and_ties(ext):
    pts = room_points(3, 8192, seed=1)
    pts[1, 4096:] = pts[1, :4096]          # CropPad-style duplication: exact ties everywhere
    pts[2] = np.round(pts[2] * 20) / 20     # coarse lattice: many equal distances
    want = oracle.farthest_point_sample(pts, 2048)
    got = ext.fps_cuda.farthest_point_sample(cu(pts), 2048).cpu().numpy()
    assert np.array_equal(got, want)
    # all points identical: the reference keeps returning index 0
    same = np.ones((1, 256, 3), np.float32)
    assert (ext.fps_cuda.farthest_point_sample(cu(same), 16).cpu().numpy() == 0).all()


@pytest.mark.parametrize('n', [100, 300, 1000, 1024, 2048, 3000, 4096, 8192])
@pytest.mark.parametrize('q', [4, 20])
def test_fps_tie_order_vs_oracle(ext, n, q):
    """Exact ties at every launch geometry (N = 1024 / 2048 / 3000 were the round-1 hole), M = N / 2."""
    rng = np.random.RandomState(n + q)
    pts = np.round(rng.rand(2, n, 3).astype(np.float32) * 2.0 * q) / q
    pts[:, n // 2:n // 2 * 2] = pts[:, :n // 2]
    want = oracle.farthest_point_sample(pts, n // 2)
    got = ext.fps_cuda.farthest_point_sample(cu(pts), n // 2).cpu().numpy()
    assert np.array_equal(got, want)


@pytest.mark.parametrize('b,n,m,q', [(2, 20000, 1500, 8), (1, 70000, 600, 0), (3, 9001, 1000, 20), (1, 140000, 300, 4)])
def test_fps_large_clouds_vs_oracle(ext, b, n, m, q):
    """Whole-scene sizes (N > 8192: the slab kernel with the sorted cloud in a global workspace), random and lattice
    clouds with exact ties, one and several slabs' worth of points per lane."""
    rng = np.random.RandomState(n + q)
    pts = (rng.rand(b, n, 3) * np.array([6.0, 8.0, 2.7])).astype(np.float32)
    if q:
        pts = (np.round(pts * q) / q).astype(np.float32)
        pts[:, n // 2:n // 2 * 2] = pts[:, :n // 2]
    want = oracle.farthest_point_sample(pts, m)
    got = ext.fps_cuda.farthest_point_sample(cu(pts), m).cpu().numpy()
    assert np.array_equal(got, want)


def test_fps_errors(ext):
    with pytest.raises(RuntimeError):
        ext.fps_cuda.farthest_point_sample(torch.zeros(1, 8, 4).cuda(), 2)
    with pytest.raises(RuntimeError):
        ext.fps_cuda.farthest_point_sample(torch.zeros(1, 8, 3).cuda(), 9)
    with pytest.raises(RuntimeError):
        ext.fps_cuda.farthest_point_sample(torch.zeros(1, 8, 3).cuda(), 0)
    with pytest.raises(RuntimeError):
        ext.fps_cuda.farthest_point_sample(torch.zeros(1, 8, 3), 2)  # CPU tensor


# ---------------------------------------------------------------- ball query
@pytest.mark.parametrize('b,n1,n2,r,k,dtype', [
    (2, 64, 128, 0.1, 32, np.float64), (3, 65, 129, 0.1, 32, np.float64), (3, 65, 129, 10.0, 32, np.float64),
    (3, 65, 129, 0.1, 32, np.float32), (3, 65, 129, 10.0, 32, np.float32), (4, 512, 1024, 0.1, 64, np.float32),
    (2, 100, 9001, 0.5, 7, np.float32), (1, 33, 20011, 0.3, 40, np.float64), (1, 5, 3, 1.0, 4, np.float32)])
def test_ball_query(ext, b, n1, n2, r, k, dtype):
    np.random.seed(0)
    key = np.random.randn(b, n2, 3).astype(dtype)
    query = np.stack([p[np.random.choice(n2, n1, replace=n1 > n2)] for p in key])
    want_i, want_d = oracle.ball_query(query, key, r, k, with_distance=True)
    got = ext.ball_query_cuda.ball_query(cu(query), cu(key), r, k)
    assert np.array_equal(got.cpu().numpy(), want_i)
    gi, gd = ext.ball_query_distance_cuda.ball_query_distance(cu(query), cu(key), r, k)
    assert np.array_equal(gi.cpu().numpy(), want_i)
    assert np.array_equal(gd.cpu().numpy(), want_d)


def test_ball_query_room_levels(ext):
    """The four SA levels of the model on the synthetic room, fp32, plus empty rows."""
    pts = room_points(2, 8192, seed=3)
    for n1, r in [(2048, 0.1), (512, 0.2), (128, 0.4), (32, 0.8)]:
        q = pts[:, :n1].copy()
        want = oracle.ball_query(q, pts, r, 32)
        got = ext.ball_query_cuda.ball_query(cu(q), cu(pts), r, 32).cpu().numpy()
        assert np.array_equal(got, want)
    far = pts[:, :16] + 100.0
    got = ext.ball_query_cuda.ball_query(cu(far), cu(pts), 0.1, 32).cpu().numpy()
    assert (got == -1).all()


def test_ball_query_threshold_is_strict_and_fp32(ext):
    # keys at exactly r (in fp32 arithmetic) must be excluded; just inside must be included
    r = np.float32(0.1)
    key = np.zeros((1, 4, 3), np.float32)
    key[0, 1, 0] = r
    key[0, 2, 0] = np.nextafter(r, np.float32(0))
    key[0, 3, 1] = np.nextafter(r, np.float32(1))
    q = np.zeros((1, 1, 3), np.float32)
    want = oracle.ball_query(q, key, float(r), 4)
    got = ext.ball_query_cuda.ball_query(cu(q), cu(key), float(r), 4).cpu().numpy()
    assert np.array_equal(got, want)


# ---------------------------------------------------------------- 3-NN
@pytest.mark.parametrize('b,n1,n2,dtype', [
    (2, 512, 1024, np.float32), (3, 513, 1025, np.float32), (3, 31, 63, np.float32), (2, 8192, 2048, np.float32),
    (2, 128, 32, np.float32), (1, 7, 3, np.float32), (2, 300, 9000, np.float64), (1, 100, 10000, np.float32)])
def test_knn_distance(ext, b, n1, n2, dtype):
    torch.manual_seed(0)
    q = torch.randn(b, n1, 3).numpy().astype(dtype)
    k = torch.randn(b, n2, 3).numpy().astype(dtype)
    want_i, want_d = oracle.knn_distance(q, k, 3)
    gi, gd = ext.knn_distance_cuda.knn_distance(cu(q), cu(k), 3)
    assert np.array_equal(gi.cpu().numpy(), want_i)
    assert np.array_equal(gd.cpu().numpy(), want_d)


def test_knn_ties_and_errors(ext):
    key = np.zeros((1, 70, 3), np.float32)
    key[0, 35:] = 1.0
    q = np.zeros((1, 2, 3), np.float32)
    want_i, _ = oracle.knn_distance(q, key, 3)
    gi, _ = ext.knn_distance_cuda.knn_distance(cu(q), cu(key), 3)
    assert np.array_equal(gi.cpu().numpy(), want_i) and want_i[0, 0].tolist() == [0, 1, 2]
    pts = np.round(room_points(1, 2048, seed=5) * 10) / 10  # lattice: many exact ties
    want_i, want_d = oracle.knn_distance(pts, pts[:, :512], 3)
    gi, gd = ext.knn_distance_cuda.knn_distance(cu(pts), cu(pts[:, :512]), 3)
    assert np.array_equal(gi.cpu().numpy(), want_i) and np.array_equal(gd.cpu().numpy(), want_d)
    with pytest.raises(RuntimeError):
        ext.knn_distance_cuda.knn_distance(cu(q), cu(key), 4)
    with pytest.raises(RuntimeError):
        ext.knn_distance_cuda.knn_distance(cu(q), cu(key[:, :2]), 3)


# ---------------------------------------------------------------- group points
@pytest.mark.parametrize('b,c,n1,n2,k,dtype', [
    (2, 3, 512, 128, 32, torch.float32), (5, 64, 513, 129, 33, torch.float32), (4, 32, 1024, 512, 64, torch.float32),
    (2, 67, 8192, 2048, 32, torch.float32), (2, 5, 100, 7, 3, torch.float64)])
def test_group_points(ext, b, c, n1, n2, k, dtype):
    torch.manual_seed(0)
    x = torch.randn(b, c, n1, dtype=dtype)
    idx = torch.randint(0, n1, [b, n2, k])
    want = oracle.group_points_forward(x.numpy(), idx.numpy())
    got = ext.group_points_cuda.group_points_forward(x.cuda(), idx.cuda())
    assert np.array_equal(got.cpu().numpy(), want)
    g = torch.randn(b, c, n2, k, dtype=dtype)
    want_g = oracle.group_points_backward(g.numpy(), idx.numpy(), n1)
    got_g = ext.group_points_cuda.group_points_backward(g.cuda(), idx.cuda(), n1)
    np.testing.assert_allclose(got_g.cpu().numpy(), want_g, rtol=1e-5, atol=1e-5)


def test_group_points_strided_input_and_autograd(ext):
    """MVPNet3D gathers from a permuted view (mvpnet_3d.py:106-109): strides must be honoured."""
    from mvpnet_b200.ops import group_points
    torch.manual_seed(1)
    xyz = torch.randn(2, 5 * 12 * 16, 3).cuda()            # (b, P, 3)
    view = xyz.permute(0, 2, 1)                              # (b, 3, P), non-contiguous
    idx = torch.randint(0, view.size(2), [2, 64, 3]).cuda()
    got = group_points(view, idx)
    want = oracle.group_points_forward(view.contiguous().cpu().numpy(), idx.cpu().numpy())
    assert np.array_equal(got.cpu().numpy(), want)
    x = torch.randn(2, 8, 50, device='cuda', requires_grad=True)
    idx = torch.randint(0, 50, [2, 10, 4]).cuda()
    y = group_points(x, idx)
    y.backward(torch.ones_like(y))
    ref = torch.zeros(2, 8, 50)
    ref.scatter_add_(2, idx.cpu().reshape(2, 1, 40).expand(2, 8, 40), torch.ones(2, 8, 40))
    np.testing.assert_allclose(x.grad.cpu().numpy(), ref.numpy(), rtol=1e-6)


def test_group_points_negative_index_is_defined(ext):
    """ball_query emits -1 rows for empty balls; the reference device-asserts on them
    (group_points_kernel.cu:85).  Here: gathers 0, skips in backward, and is counted."""
    ext.index_errors_fetch_and_clear()
    x = torch.arange(12, dtype=torch.float32).reshape(1, 2, 6).cuda() + 1
    idx = torch.tensor([[[0, -1, 5, 6]]]).cuda()
    out = ext.group_points_cuda.group_points_forward(x, idx).cpu().numpy()
    assert out[0, 0, 0].tolist() == [1.0, 0.0, 6.0, 0.0]
    assert ext.index_errors_fetch_and_clear() == 2
    g = ext.group_points_cuda.group_points_backward(torch.ones(1, 2, 1, 4).cuda(), idx, 6).cpu().numpy()
    assert g[0, 0].tolist() == [1, 0, 0, 0, 0, 1]
    assert ext.index_errors_fetch_and_clear() == 2


# ---------------------------------------------------------------- interpolate
@pytest.mark.parametrize('b,c,m,n,dtype', [
    (2, 64, 128, 512, torch.float64), (3, 65, 129, 513, torch.float64), (2, 64, 256, 1024, torch.float32),
    (2, 128, 2048, 8192, torch.float32), (2, 512, 32, 128, torch.float32)])
def test_interpolate(ext, b, c, m, n, dtype):
    torch.manual_seed(0)
    x = torch.randn(b, c, m, dtype=dtype)
    idx = torch.randint(0, m, [b, n, 3])
    w = torch.rand(b, n, 3, dtype=dtype)
    want = oracle.interpolate_forward(x.numpy(), idx.numpy(), w.numpy())
    got = ext.interpolate_cuda.interpolate_forward(x.cuda(), idx.cuda(), w.cuda())
    assert np.array_equal(got.cpu().numpy(), want)
    g = torch.randn(b, c, n, dtype=dtype)
    want_g = oracle.interpolate_backward(g.numpy(), idx.numpy(), w.numpy(), m)
    got_g = ext.interpolate_cuda.interpolate_backward(g.cuda(), idx.cuda(), w.cuda(), m)
    tol = 1e-4 if dtype == torch.float32 else 1e-10
    np.testing.assert_allclose(got_g.cpu().numpy(), want_g, rtol=tol, atol=tol)


def test_interpolate_autograd_matches_torch(ext):
    from mvpnet_b200.ops import feature_interpolate
    torch.manual_seed(2)
    x = torch.randn(2, 16, 40, dtype=torch.float64, device='cuda', requires_grad=True)
    idx = torch.randint(0, 40, [2, 90, 3]).cuda()
    w = torch.rand(2, 90, 3, dtype=torch.float64).cuda()
    y = feature_interpolate(x, idx, w)
    x2 = x.detach().clone().requires_grad_(True)
    gathered = x2.unsqueeze(2).expand(2, 16, 90, 40).gather(3, idx.unsqueeze(1).expand(2, 16, 90, 3))
    y2 = (gathered * w.unsqueeze(1)).sum(3)
    go = torch.randn_like(y)
    y.backward(go)
    y2.backward(go)
    np.testing.assert_allclose(y.detach().cpu().numpy(), y2.detach().cpu().numpy(), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(x.grad.cpu().numpy(), x2.grad.cpu().numpy(), rtol=1e-10, atol=1e-12)


# ---------------------------------------------------------------- full-size properties (BASELINE sizes, B=32)
def test_full_size_properties(ext):
    b, n = 32, 8192
    pts = cu(room_points(b, n, seed=7))
    idx = ext.fps_cuda.farthest_point_sample(pts, 2048)
    s = torch.sort(idx, dim=1)[0]
    assert (idx[:, 0] == 0).all() and (s[:, 1:] != s[:, :-1]).all()          # distinct, starts at 0
    cent = torch.gather(pts, 1, idx.unsqueeze(-1).expand(b, 2048, 3)).contiguous()
    bq = ext.ball_query_cuda.ball_query(cent, pts, 0.1, 32)
    assert (bq >= 0).all() and (bq < n).all()
    d = (torch.gather(pts, 1, bq.reshape(b, -1, 1).expand(b, 2048 * 32, 3)).reshape(b, 2048, 32, 3)
         - cent.unsqueeze(2)).pow(2).sum(-1)
    assert (d < 0.1 * 0.1 + 1e-6).all()                                       # inside the ball
    first = bq[:, :, :1]
    inc = (bq[:, :, 1:] > bq[:, :, :-1]) | (bq[:, :, 1:] == first)            # ascending, then first-hit pad
    assert inc.all()
    ki, kd = ext.knn_distance_cuda.knn_distance(pts, cent, 3)
    assert (kd[:, :, 1:] >= kd[:, :, :-1]).all()
    # every centroid is its own nearest neighbour at distance 0
    own_i, own_d = ext.knn_distance_cuda.knn_distance(cent, cent, 3)
    assert (own_d[:, :, 0] == 0).all()
    # oracle on one cloud of the batch
    c = 17
    assert np.array_equal(idx[c].cpu().numpy(), oracle.farthest_point_sample(pts[c:c + 1].cpu().numpy(), 2048)[0])
