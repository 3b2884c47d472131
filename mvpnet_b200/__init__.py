"""mvpnet_b200 — B200-native (sm_100a) implementation of the MVPNet hot path.

Layout
  csrc/                CUDA kernels + C ABI (include/mvpnet_b200.h) + torch shim
  libmvpnet_b200.so    built by `python -m mvpnet_b200.build` (nvcc, sm_100a)
  _ext*.so             torch extension exposing the reference's six extension modules
  ops/                 upper face: same functions / autograd.Functions as `mvpnet.ops`
  compat.py            registers the extension under the reference's module names

There is no CPU fallback: importing `mvpnet_b200.ext` without the built extension raises.
"""
import importlib

__all__ = ['ext', 'load_ext']
_EXT = None


def load_ext():
    """Import the native extension; fail loudly (never fall back) when it is missing."""
    global _EXT
    if _EXT is None:
        import torch  # noqa: F401  (libtorch must be loaded before the extension)
        try:
            _EXT = importlib.import_module('mvpnet_b200._ext')
        except ImportError as e:  # pragma: no cover
            raise ImportError(
                'mvpnet_b200: native extension not built or not loadable (%s). '
                'Run `python -m mvpnet_b200.build` (needs nvcc for sm_100a); there is no CPU/eager fallback.' % e
            ) from e
    return _EXT


def __getattr__(name):
    if name == 'ext':
        return load_ext()
    raise AttributeError(name)
