"""Drop-in registration for the reference code base.

The reference's Python wrappers import their native modules as `from . import fps_cuda`
(mvpnet/ops/fps.py:2, ball_query.py:2-3, group_points.py:2, knn_distance.py:2, interpolate.py:2).
`install()` publishes this package's extension sub-modules under exactly those names, so the
reference's unmodified `mvpnet.ops.*`, `mvpnet.models.*`, `train_mvpnet_3d.py`, `test_mvpnet_3d.py`
run on the sm_100a kernels without building `mvpnet/ops/setup.py`."""
import importlib
import sys
import types

EXT_MODULES = ('fps_cuda', 'ball_query_cuda', 'ball_query_distance_cuda', 'group_points_cuda',
               'knn_distance_cuda', 'interpolate_cuda')


def install(modules=None, reference_root=None):
    """Register `modules` (default: the CUDA extension) as mvpnet.ops.<name>.

    modules: optional {name: module}; used by the tests to put the CPU oracle behind the reference.
    reference_root: optional path of a reference checkout to put on sys.path.
    """
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    if modules is None:
        from . import load_ext
        ext = load_ext()
        modules = {name: getattr(ext, name) for name in EXT_MODULES}
    for name in EXT_MODULES:
        sys.modules['mvpnet.ops.' + name] = modules[name]
    try:
        pkg = importlib.import_module('mvpnet.ops')
    except ImportError:
        pkg = None
    if pkg is not None:
        for name in EXT_MODULES:
            setattr(pkg, name, modules[name])
    return modules


def uninstall():
    for name in EXT_MODULES:
        sys.modules.pop('mvpnet.ops.' + name, None)
    for name in [m for m in sys.modules if m == 'mvpnet' or m.startswith('mvpnet.') or m == 'common' or m.startswith('common.')]:
        sys.modules.pop(name, None)
