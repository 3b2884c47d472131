"""CPU: the C oracle against the independent numpy/torch oracles of the reference's own ops
tests (mvpnet/ops/tests/test_{fps,ball_query,group_points,interpolate,knn_distance}.py), same
seeds, same shapes, same dtypes.  These are the only pinned behaviours the reference has for
the six extension functions."""
import numpy as np
import pytest
import torch

import oracle


# ---- test_fps.py:7-37 restated: greedy FPS, np.argmax => lowest index on ties ---------------
def fps_np(points, m):
    out = []
    for p in points:
        idx, cur, d2s = [0], 0, None
        for _ in range(1, m):
            d = np.square(p - p[cur][None]).sum(1)
            d2s = d if d2s is None else np.minimum(d, d2s)
            cur = int(np.argmax(d2s))
            idx.append(cur)
        out.append(idx)
    return np.asarray(out)


@pytest.mark.parametrize('b,c,n,m,transpose', [
    (2, 3, 1024, 128, True), (2, 2, 1024, 128, True), (3, 3, 1025, 129, True),
    (3, 3, 1025, 129, False), (4, 3, 1024, 512, True), (2, 3, 8192, 2048, True)])
def test_fps(b, c, n, m, transpose):
    np.random.seed(0)
    pts = np.random.rand(b, c, n) if transpose else np.random.rand(b, n, c)
    bnd = np.transpose(pts, [0, 2, 1]) if transpose else pts
    assert np.array_equal(oracle.farthest_point_sample(bnd, m), fps_np(bnd, m))
    # fp32 inputs too (what the models feed)
    b32 = bnd.astype(np.float32)
    got = oracle.farthest_point_sample(b32, m)
    assert got.shape == (b, m) and got.dtype == np.int64 and (got[:, 0] == 0).all()
    assert all(len(set(r.tolist())) == m for r in got)


def test_fps_tie_rule_and_duplicates():
    """Reference tie rule (fps_kernel.cu:95-129).  Thread t = j mod BLOCK keeps its first strict
    max; the shared-memory tree (`if (dist1 < dist2)` at offsets BLOCK/2..1) keeps the LOWER
    position on ties, and because values migrate towards position 0 the survivor among equal
    maxima is the thread with the smallest BIT-REVERSED id (even t beats odd t at the last level,
    t%4==0 beats t%4==2 one level up, ...), then the smallest j within that thread.
    Duplicated points matter: CropPad pads by duplication (transforms.py:122-125)."""
    # 4 distinct corners repeated; N=1024 => BLOCK=512: j and j+512 share a thread.
    base = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], np.float32)
    pts = np.tile(base, (256, 1))[None]  # (1,1024,3)
    got = oracle.farthest_point_sample(pts, 6)[0]
    # after 0: farthest is corner 3 (dist 2) -> candidates j%4==3, smallest (j mod 512) then j => 3
    assert got[0] == 0 and got[1] == 3
    # then corners 1 and 2 tie at dist 1 -> thread 2 (bit-reversed 010..0) beats thread 1 (10..0)
    assert got[2] == 2 and got[3] == 1
    # all remaining distances are zero -> reference keeps returning the previous index
    assert got[4] == 1 and got[5] == 1
    assert oracle.ref_block_size(1024) == 512 and oracle.ref_block_size(128) == 128
    assert oracle.ref_block_size(1025) == 512 and oracle.ref_block_size(129) == 128
    assert oracle.ref_block_size(8) == 16 and oracle.ref_block_size(8192) == 512


def test_fps_errors():
    with pytest.raises(RuntimeError):
        oracle.farthest_point_sample(np.zeros((1, 8, 4)), 2)
    with pytest.raises(RuntimeError):
        oracle.farthest_point_sample(np.zeros((1, 8, 3)), 9)


# ---- test_ball_query.py:16-41,71-98 restated ------------------------------------------------
def ball_query_np(query, key, radius, k):
    idx_all, dist_all = [], []
    for q, kk in zip(query, key):
        idx = np.full([q.shape[0], k], -1, np.int64)
        dist = np.full([q.shape[0], k], -1.0, q.dtype)
        for i in range(q.shape[0]):
            d = np.square(kk - q[i][None]).sum(1)
            nb = np.nonzero(d < radius ** 2)[0]
            if nb.size == 0:
                continue
            n = min(nb.size, k)
            idx[i, :n] = nb[:n]
            idx[i, n:] = nb[0]
            dist[i, :n] = d[nb[:n]]
        idx_all.append(idx)
        dist_all.append(dist)
    return np.asarray(idx_all), np.asarray(dist_all)


@pytest.mark.parametrize('b,n1,n2,r,k', [
    (2, 64, 128, 0.1, 32), (3, 65, 129, 0.1, 32), (3, 65, 129, 10.0, 32), (4, 512, 1024, 0.1, 64)])
def test_ball_query(b, n1, n2, r, k):
    np.random.seed(0)
    key = np.random.randn(b, 3, n2)
    query = np.array([p[:, np.random.choice(n2, n1, replace=False)] for p in key])
    key, query = key.transpose(0, 2, 1), query.transpose(0, 2, 1)
    want_i, want_d = ball_query_np(query, key, r, k)
    got_i = oracle.ball_query(query, key, r, k)
    got_i2, got_d = oracle.ball_query(query, key, r, k, with_distance=True)
    assert np.array_equal(got_i, want_i) and np.array_equal(got_i2, want_i)
    np.testing.assert_allclose(got_d, want_d, rtol=1e-12, atol=0)


def test_ball_query_no_hit_rows_stay_minus_one():
    key = np.zeros((1, 4, 3), np.float32)
    query = np.full((1, 2, 3), 5.0, np.float32)
    idx, dist = oracle.ball_query(query, key, 0.1, 8, with_distance=True)
    assert (idx == -1).all() and (dist == -1).all()


# ---- test_group_points.py:6-44 restated ------------------------------------------------------
@pytest.mark.parametrize('b,c,n1,n2,k', [(2, 3, 512, 128, 32), (5, 64, 513, 129, 33)])
def test_group_points(b, c, n1, n2, k):
    torch.manual_seed(0)
    x = torch.randn(b, c, n1, requires_grad=True)
    idx = torch.randint(0, n1, [b, n2, k])
    want = x.unsqueeze(2).expand(b, c, n2, n1).gather(3, idx.unsqueeze(1).expand(b, c, n2, k))
    got = oracle.group_points_forward(x.detach().numpy(), idx.numpy())
    assert np.array_equal(got, want.detach().numpy())
    want.backward(torch.ones_like(want))
    gin = oracle.group_points_backward(np.ones((b, c, n2, k), np.float32), idx.numpy(), n1)
    np.testing.assert_allclose(gin, x.grad.numpy(), rtol=1e-6)


# ---- test_interpolate.py:6-64 restated -------------------------------------------------------
@pytest.mark.parametrize('b,c,m,n', [(2, 64, 128, 512), (3, 65, 129, 513)])
def test_interpolate(b, c, m, n):
    torch.manual_seed(0)
    x = torch.randn(b, c, m, dtype=torch.float64, requires_grad=True)
    idx = torch.randint(0, m, [b, n, 3])
    w = torch.rand(b, n, 3, dtype=torch.float64)
    g = x.unsqueeze(2).expand(b, c, n, m).gather(3, idx.unsqueeze(1).expand(b, c, n, 3))
    want = (g * w.unsqueeze(1)).sum(3)
    got = oracle.interpolate_forward(x.detach().numpy(), idx.numpy(), w.numpy())
    np.testing.assert_allclose(got, want.detach().numpy(), rtol=1e-12, atol=1e-14)
    go = torch.randn(b, c, n, dtype=torch.float64)
    want.backward(go)
    gin = oracle.interpolate_backward(go.numpy(), idx.numpy(), w.numpy(), m)
    np.testing.assert_allclose(gin, x.grad.numpy(), rtol=1e-10, atol=1e-12)


# ---- test_knn_distance.py:7-54 restated ------------------------------------------------------
@pytest.mark.parametrize('b,n1,n2', [(2, 512, 1024), (3, 513, 1025), (3, 31, 63)])
def test_knn_distance(b, n1, n2):
    torch.manual_seed(0)
    q = torch.randn(b, n1, 3)
    k = torch.randn(b, n2, 3)
    d = ((q.unsqueeze(2) - k.unsqueeze(1)) ** 2).sum(3)
    want_d, want_i = torch.topk(d, 3, dim=2, largest=False, sorted=True)
    got_i, got_d = oracle.knn_distance(q.numpy(), k.numpy(), 3)
    assert np.array_equal(got_i, want_i.numpy())
    np.testing.assert_allclose(got_d, want_d.numpy(), atol=1e-6)


def test_knn_ties_take_lowest_index():
    key = np.zeros((1, 6, 3), np.float32)
    key[0, 3:] = 1.0
    idx, dist = oracle.knn_distance(np.zeros((1, 1, 3), np.float32), key, 3)
    assert idx.tolist() == [[[0, 1, 2]]] and (dist == 0).all()
    with pytest.raises(RuntimeError):
        oracle.knn_distance(np.zeros((1, 1, 3), np.float32), key[:, :2], 3)
