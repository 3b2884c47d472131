"""CPU oracle for the MVPNet hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  It wraps ``oracle/mvp_oracle.c`` (a plain-C
restatement of the reference's CUDA kernels and data-side numpy code; every function there cites
the reference file:line it follows) behind numpy arrays, and offers ``ext_modules()``: six
objects shaped like the reference's pybind11 extension modules (``mvpnet/ops/cuda/*.cpp``) so
that the reference's *unmodified* Python (``mvpnet.ops.*``, ``mvpnet.models.*``) can be run on
CPU on top of it to generate golden vectors and to time the CPU baseline.

Pinning status: pinned against the reference's own ops-test oracles (re-run in
``tests/test_oracle_ops.py``) and against fixtures made by importing the reference's Python
(``tests/golden/make_golden.py``).  The 2D->3D k-NN restates scikit-learn (third party, unpinned
upstream); pinned against scikit-learn 1.9.0 in ``tests/test_oracle_unproject.py``.
"""
import ctypes
import os
import subprocess
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i64 = ctypes.c_int64
_p = ctypes.c_void_p


def build(force=False):
    """Compile the C oracle with gcc (Makefile next to this file)."""
    if force:
        subprocess.run(['make', '-C', _HERE, 'clean'], check=True, capture_output=True)
    subprocess.run(['make', '-C', _HERE], check=True, capture_output=True)


def _cpu_has_fma():
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.startswith('flags'):
                    return ' fma ' in line + ' '
    except OSError:
        pass
    return False


def lib():
    global _LIB
    if _LIB is None:
        name = 'libmvp_oracle_fma.so' if _cpu_has_fma() else 'libmvp_oracle.so'
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.mvpo_ref_block_size.argtypes = [_i64, ctypes.c_int]
        _LIB.mvpo_ref_block_size.restype = ctypes.c_int
    return _LIB


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _sfx(dtype):
    if dtype == np.float32:
        return '_f32'
    if dtype == np.float64:
        return '_f64'
    raise TypeError('oracle supports float32/float64, got %s' % dtype)


def _call(name, *args):
    fn = getattr(lib(), name)
    fn.restype = ctypes.c_int
    rc = fn(*args)
    if rc != 0:
        raise RuntimeError('%s failed with code %d' % (name, rc))


def _c(a, dtype=None):
    return np.ascontiguousarray(a, dtype=dtype)


# ---------------------------------------------------------------------------------------------
# numpy-level functions (shapes exactly as the reference extension functions take them)
# ---------------------------------------------------------------------------------------------
def farthest_point_sample(points, num_centroids):
    """points (B, N, D in {2,3}) -> int64 (B, M).  fps.cpp:7-13 / fps_kernel.cu:144-180."""
    points = _c(points)
    B, N, D = points.shape
    if D not in (2, 3):
        raise RuntimeError('Only support dim=2 or dim=3')
    if not (num_centroids > 0 and N >= num_centroids):
        raise RuntimeError('num_centroids must be in (0, num_points]')
    out = np.zeros((B, num_centroids), dtype=np.int64)
    _call('mvpo_fps' + _sfx(points.dtype), _ptr(points), _i64(B), _i64(N), _i64(D),
          _i64(num_centroids), _ptr(out))
    return out


def ball_query(query, key, radius, max_neighbors, with_distance=False):
    """query (B,N1,3), key (B,N2,3) -> int64 (B,N1,K) [, dist (B,N1,K)]."""
    query, key = _c(query), _c(key)
    if query.shape[2] != 3 or key.shape[2] != 3:
        raise RuntimeError('xyz must have 3 channels')
    B, N1, _ = query.shape
    N2 = key.shape[1]
    index = np.empty((B, N1, max_neighbors), dtype=np.int64)
    dist = np.empty((B, N1, max_neighbors), dtype=query.dtype) if with_distance else None
    _call('mvpo_ball_query' + _sfx(query.dtype), _ptr(query), _ptr(key), _i64(B), _i64(N1),
          _i64(N2), ctypes.c_float(radius), _i64(max_neighbors), _ptr(index), _ptr(dist))
    return (index, dist) if with_distance else index


def knn_distance(query, key, k):
    query, key = _c(query), _c(key)
    if k != 3:
        raise RuntimeError('Only support 3-NN.')
    B, N1, _ = query.shape
    N2 = key.shape[1]
    if N2 < k:
        raise RuntimeError('num_key must be >= k')
    index = np.empty((B, N1, 3), dtype=np.int64)
    dist = np.empty((B, N1, 3), dtype=query.dtype)
    _call('mvpo_knn3' + _sfx(query.dtype), _ptr(query), _ptr(key), _i64(B), _i64(N1), _i64(N2),
          _ptr(index), _ptr(dist))
    return index, dist


def group_points_forward(inp, index):
    inp, index = _c(inp), _c(index, np.int64)
    B, C, N1 = inp.shape
    _, N2, K = index.shape
    out = np.empty((B, C, N2, K), dtype=inp.dtype)
    _call('mvpo_group_points_fwd' + _sfx(inp.dtype), _ptr(inp), _ptr(index), _i64(B), _i64(C),
          _i64(N1), _i64(N2), _i64(K), _ptr(out))
    return out


def group_points_backward(grad_out, index, num_points):
    grad_out, index = _c(grad_out), _c(index, np.int64)
    B, C, N2, K = grad_out.shape
    gin = np.empty((B, C, num_points), dtype=grad_out.dtype)
    _call('mvpo_group_points_bwd' + _sfx(grad_out.dtype), _ptr(grad_out), _ptr(index), _i64(B),
          _i64(C), _i64(num_points), _i64(N2), _i64(K), _ptr(gin))
    return gin


def interpolate_forward(inp, index, weight):
    inp, index = _c(inp), _c(index, np.int64)
    weight = _c(weight, inp.dtype)
    B, C, M = inp.shape
    N = index.shape[1]
    if index.shape[2] != 3:
        raise RuntimeError('k must be 3')
    out = np.empty((B, C, N), dtype=inp.dtype)
    _call('mvpo_interpolate_fwd' + _sfx(inp.dtype), _ptr(inp), _ptr(index), _ptr(weight), _i64(B),
          _i64(C), _i64(M), _i64(N), _ptr(out))
    return out


def interpolate_backward(grad_out, index, weight, num_inst):
    grad_out, index = _c(grad_out), _c(index, np.int64)
    weight = _c(weight, grad_out.dtype)
    B, C, N = grad_out.shape
    gin = np.empty((B, C, num_inst), dtype=grad_out.dtype)
    _call('mvpo_interpolate_bwd' + _sfx(grad_out.dtype), _ptr(grad_out), _ptr(index), _ptr(weight),
          _i64(B), _i64(C), _i64(num_inst), _i64(N), _ptr(gin))
    return gin


def unproject(depth, cam_inv, pose, chunk_box=None):
    """depth (nv,h,w) f32 metres; cam_inv (nv,3,3) f32 = np.linalg.inv(K[:3,:3]); pose (nv,4,4) f32.

    Returns xyz64 (nv,h,w,3) float64, xyz32 float32, mask (nv,h,w) bool.
    """
    depth = _c(depth, np.float32)
    nv, h, w = depth.shape
    cam_inv = _c(np.broadcast_to(np.asarray(cam_inv, np.float32), (nv, 3, 3)))
    pose = _c(np.broadcast_to(np.asarray(pose, np.float32), (nv, 4, 4)))
    box = None if chunk_box is None else _c(np.asarray(chunk_box, np.float64)[:4])
    xyz64 = np.empty((nv, h, w, 3), np.float64)
    xyz32 = np.empty((nv, h, w, 3), np.float32)
    mask = np.empty((nv, h, w), np.uint8)
    _call('mvpo_unproject', _ptr(depth), _ptr(cam_inv), _ptr(pose), _i64(nv), _i64(h), _i64(w),
          _ptr(box), _ptr(xyz64), _ptr(xyz32), _ptr(mask))
    return xyz64, xyz32, mask.astype(bool)


def knn_pixels(query, pix_xyz, mask, k):
    """query (nq,3), pix_xyz (P,3) float64, mask (P,) -> flat pixel ids int64 (nq,k), d^2 (nq,k)."""
    query = _c(query, np.float64)
    pix_xyz = _c(np.asarray(pix_xyz, np.float64).reshape(-1, 3))
    mask = _c(np.asarray(mask).reshape(-1), np.uint8)
    nq, P = query.shape[0], pix_xyz.shape[0]
    index = np.empty((nq, k), np.int64)
    dist = np.empty((nq, k), np.float64)
    _call('mvpo_knn_pixels', _ptr(query), _ptr(pix_xyz), _ptr(mask), _i64(nq), _i64(P), _i64(k),
          _ptr(index), _ptr(dist))
    return index, dist


def ref_block_size(n, floor_size=16):
    return lib().mvpo_ref_block_size(int(n), int(floor_size))


# ---------------------------------------------------------------------------------------------
# torch-facing stand-ins for the reference's six pybind11 extension modules (CPU tensors)
# ---------------------------------------------------------------------------------------------
def ext_modules():
    """Return {name: module-like} for fps_cuda, ball_query_cuda, ball_query_distance_cuda,
    group_points_cuda, knn_distance_cuda, interpolate_cuda operating on CPU torch tensors."""
    import torch

    def t2n(t):
        return t.detach().cpu().contiguous().numpy()

    def n2t(a):
        return torch.from_numpy(a)

    m = {}
    fps = types.ModuleType('fps_cuda')
    fps.farthest_point_sample = lambda points, m_: n2t(farthest_point_sample(t2n(points), int(m_)))
    m['fps_cuda'] = fps

    bq = types.ModuleType('ball_query_cuda')
    bq.ball_query = lambda q, k, r, K: n2t(ball_query(t2n(q), t2n(k), float(r), int(K)))
    m['ball_query_cuda'] = bq

    bqd = types.ModuleType('ball_query_distance_cuda')

    def _bqd(q, k, r, K):
        i, d = ball_query(t2n(q), t2n(k), float(r), int(K), with_distance=True)
        return [n2t(i), n2t(d)]
    bqd.ball_query_distance = _bqd
    m['ball_query_distance_cuda'] = bqd

    gp = types.ModuleType('group_points_cuda')
    gp.group_points_forward = lambda x, i: n2t(group_points_forward(t2n(x), t2n(i)))
    gp.group_points_backward = lambda g, i, n: n2t(group_points_backward(t2n(g), t2n(i), int(n)))
    m['group_points_cuda'] = gp

    knn = types.ModuleType('knn_distance_cuda')

    def _knn(q, k, kk):
        i, d = knn_distance(t2n(q), t2n(k), int(kk))
        return [n2t(i), n2t(d)]
    knn.knn_distance = _knn
    m['knn_distance_cuda'] = knn

    it = types.ModuleType('interpolate_cuda')
    it.interpolate_forward = lambda x, i, w: n2t(interpolate_forward(t2n(x), t2n(i), t2n(w)))
    it.interpolate_backward = lambda g, i, w, n: n2t(interpolate_backward(t2n(g), t2n(i), t2n(w), int(n)))
    m['interpolate_cuda'] = it
    return m
