"""CPU: the C-ABI library and the torch shim load, and every symbol include/mvpnet_b200.h declares
is exported (no compute call — there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'mvpnet_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(mvp_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def built():
    from mvpnet_b200 import build
    return build.build_all()


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built[0])
    names = declared_symbols()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), 'missing export: ' + n
    lib.mvp_abi_version.restype = ctypes.c_int
    assert lib.mvp_abi_version() >= 1


def test_argument_errors_are_reported_without_a_gpu(built):
    lib = ctypes.CDLL(built[0])
    lib.mvp_last_error.restype = ctypes.c_char_p
    i64 = ctypes.c_int64
    rc = lib.mvp_fps(None, i64(1), i64(8), i64(4), i64(2), 0, None, None, None)
    assert rc == -1 and b'dim=2 or dim=3' in lib.mvp_last_error()
    rc = lib.mvp_fps(None, i64(1), i64(8), i64(3), i64(9), 0, None, None, None)
    assert rc == -1
    rc = lib.mvp_knn_distance(None, None, i64(1), i64(4), i64(8), i64(5), 0, None, None, None, None)
    assert rc == -1 and b'3-NN' in lib.mvp_last_error()
    rc = lib.mvp_knn_distance(None, None, i64(1), i64(4), i64(2), i64(3), 0, None, None, None, None)
    assert rc == -1
    lib.mvp_fps_workspace_bytes.restype = ctypes.c_int64
    assert lib.mvp_fps_workspace_bytes(i64(2), i64(8192), i64(3), i64(2048), 0) == 0
    assert lib.mvp_fps_workspace_bytes(i64(2), i64(10000), i64(3), i64(64), 1) == 2 * 10000 * 8


def test_torch_shim_exposes_reference_module_names(built):
    import mvpnet_b200
    ext = mvpnet_b200.load_ext()
    from mvpnet_b200.compat import EXT_MODULES
    for name in EXT_MODULES:
        assert hasattr(ext, name)
    assert hasattr(ext.fps_cuda, 'farthest_point_sample')
    assert hasattr(ext.ball_query_cuda, 'ball_query')
    assert hasattr(ext.ball_query_distance_cuda, 'ball_query_distance')
    assert hasattr(ext.group_points_cuda, 'group_points_forward') and hasattr(ext.group_points_cuda, 'group_points_backward')
    assert hasattr(ext.knn_distance_cuda, 'knn_distance')
    assert hasattr(ext.interpolate_cuda, 'interpolate_forward') and hasattr(ext.interpolate_cuda, 'interpolate_backward')


def test_shim_rejects_cpu_tensors_like_the_reference():
    import torch
    import mvpnet_b200
    ext = mvpnet_b200.load_ext()
    with pytest.raises(RuntimeError, match='CUDA tensor'):
        ext.fps_cuda.farthest_point_sample(torch.zeros(1, 8, 3), 2)
    with pytest.raises(RuntimeError, match='CUDA tensor'):
        ext.group_points_cuda.group_points_forward(torch.zeros(1, 2, 3), torch.zeros(1, 1, 1, dtype=torch.long))
