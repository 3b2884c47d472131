"""The exact uniform-grid ball query / 3-NN (csrc/point_grid.cu) against the exhaustive kernels of the same
library and the CPU oracle: indices, order, padding and distances must be bit-identical on every distribution,
including the ones a grid is bad at (clusters, lattices with exact ties, planes, coincident points, non-finite
coordinates -> per-cloud fallback inside one launch)."""
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


def clouds(kind, b, n, seed, dtype):
    rng = np.random.RandomState(seed)
    if kind == 'uniform':
        p = rng.rand(b, n, 3) * [1.9, 1.9, 2.5]
    elif kind == 'normal':
        p = rng.randn(b, n, 3)
    elif kind == 'clusters':            # a few very dense blobs far apart: most cells empty, some overfull
        c = rng.rand(b, 8, 3) * 4.0
        p = np.stack([c[i][rng.randint(0, 8, n)] for i in range(b)]) + rng.randn(b, n, 3) * 0.01
    elif kind == 'lattice':             # exact distance ties everywhere
        p = np.round(rng.rand(b, n, 3) * [1.9, 1.9, 2.5] * 10) / 10
    elif kind == 'plane':               # degenerate axis
        p = rng.rand(b, n, 3) * [2.0, 2.0, 0.0]
    elif kind == 'offset':              # far from the origin: coordinate rounding matters
        p = rng.rand(b, n, 3) * [1.9, 1.9, 2.5] + [512.0, -1024.0, 64.0]
    elif kind == 'coincident':
        p = np.zeros((b, n, 3)) + 0.25
    else:
        raise KeyError(kind)
    return np.ascontiguousarray(p.astype(dtype))


KINDS = ['uniform', 'normal', 'clusters', 'lattice', 'plane', 'offset', 'coincident']


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_ball_query_grid_equals_exhaustive(ext, kind, dtype):
    b, n2, n1 = 2, 8192, 1024
    key = clouds(kind, b, n2, 11, dtype)
    rng = np.random.RandomState(5)
    query = np.ascontiguousarray(key[:, rng.choice(n2, n1, replace=False)])
    query[:, ::7] += (rng.randn(b, len(range(0, n1, 7)), 3) * 0.03).astype(dtype)
    for r, k in [(0.1, 32), (0.25, 16), (0.05, 64)]:
        was = ext.set_grid_search(False)
        try:
            ei, ed = ext.ball_query_distance_cuda.ball_query_distance(cu(query), cu(key), r, k)
        finally:
            ext.set_grid_search(True)
        gi, gd = ext.ball_query_distance_cuda.ball_query_distance(cu(query), cu(key), r, k)
        gi2 = ext.ball_query_cuda.ball_query(cu(query), cu(key), r, k)
        ext.set_grid_search(was)
        assert torch.equal(gi, ei) and torch.equal(gi2, ei), (kind, r, k)
        assert torch.equal(gd, ed), (kind, r, k)
    # one cloud against the CPU oracle as well
    want_i, want_d = oracle.ball_query(query[:1, :256], key[:1], 0.1, 32, with_distance=True)
    gi, gd = ext.ball_query_distance_cuda.ball_query_distance(cu(query[:, :256]), cu(key), 0.1, 32)
    assert np.array_equal(gi[:1].cpu().numpy(), want_i) and np.array_equal(gd[:1].cpu().numpy(), want_d)


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_knn3_grid_equals_exhaustive(ext, kind, dtype):
    b, n2, n1 = 2, 2048, 4096
    key = clouds(kind, b, n2, 21, dtype)
    query = clouds(kind, b, n1, 22, dtype)
    query[:, :n2 // 2] = key[:, :n2 // 2]          # exact zero distances and shared coordinates
    was = ext.set_grid_search(False)
    try:
        ei, ed = ext.knn_distance_cuda.knn_distance(cu(query), cu(key), 3)
    finally:
        ext.set_grid_search(True)
    gi, gd = ext.knn_distance_cuda.knn_distance(cu(query), cu(key), 3)
    ext.set_grid_search(was)
    assert torch.equal(gi, ei) and torch.equal(gd, ed), kind
    want_i, want_d = oracle.knn_distance(query[:1, :512], key[:1], 3)
    assert np.array_equal(gi[:1, :512].cpu().numpy(), want_i) and np.array_equal(gd[:1, :512].cpu().numpy(), want_d)


def test_grid_nonfinite_cloud_falls_back_per_cloud(ext):
    """One cloud of the batch holds inf / nan coordinates: it is served by the exhaustive kernel inside the same
    call, the others by the grid; all rows equal the all-exhaustive result."""
    key = clouds('uniform', 3, 8192, 31, np.float32)
    key[1, 100] = np.inf
    key[1, 200, 1] = np.nan
    query = np.ascontiguousarray(key[:, :1024])
    was = ext.set_grid_search(False)
    ei = ext.ball_query_cuda.ball_query(cu(query), cu(key), 0.1, 32)
    ki, kd = ext.knn_distance_cuda.knn_distance(cu(key), cu(query), 3)
    ext.set_grid_search(True)
    gi = ext.ball_query_cuda.ball_query(cu(query), cu(key), 0.1, 32)
    gki, gkd = ext.knn_distance_cuda.knn_distance(cu(key), cu(query), 3)
    ext.set_grid_search(was)
    assert torch.equal(gi, ei)
    assert torch.equal(gki, ki) and torch.equal(torch.nan_to_num(gkd, nan=-7.0), torch.nan_to_num(kd, nan=-7.0))
