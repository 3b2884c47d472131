"""In-tree build of the native code (sm_100a only).

  libmvpnet_b200.so   nvcc, CUDA kernels + C ABI (include/mvpnet_b200.h), no torch dependency
  _ext*.so            g++, pybind11/torch shim over the C ABI (csrc/torch_ext.cpp)

`python -m mvpnet_b200.build` builds both next to this file; the built objects travel to the GPU
box with the repo snapshot.  Nothing is JIT-compiled at import time.
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libmvpnet_b200.so')
EXT = os.path.join(HERE, '_ext' + (sysconfig.get_config_var('EXT_SUFFIX') or '.so'))

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(exts):
    out = [os.path.join(ROOT, 'include', 'mvpnet_b200.h')]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith(exts):
            out.append(os.path.join(CSRC, f))
    return out


def _run(cmd, verbose):
    if verbose:
        print(' '.join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('build step failed: ' + ' '.join(cmd[:3]) + ' ...')
    return res.stdout + res.stderr


def build_lib(force=False, verbose=False, ptxas_v=False):
    srcs = _sources(('.cu', '.cuh'))
    if not force and not _newer(LIB, srcs):
        return LIB
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if ptxas_v else []) + ['-o', LIB, os.path.join(CSRC, 'lib.cu')]
    out = _run(cmd, verbose)
    if ptxas_v:
        print(out)
    return LIB


def build_ext(force=False, verbose=False):
    srcs = [os.path.join(CSRC, 'torch_ext.cpp'), os.path.join(ROOT, 'include', 'mvpnet_b200.h')]
    if not force and not _newer(EXT, srcs + [LIB]):
        return EXT
    import torch
    from torch.utils import cpp_extension as ce
    tlib = os.path.join(os.path.dirname(torch.__file__), 'lib')
    inc = []
    for p in ce.include_paths() + [sysconfig.get_paths()['include'], '/usr/local/cuda/include']:
        inc += ['-isystem', p]
    cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
    cmd = [cxx, '-O2', '-fPIC', '-shared', '-std=c++17', '-DTORCH_EXTENSION_NAME=_ext',
           '-DTORCH_API_INCLUDE_EXTENSION_H', '-D_GLIBCXX_USE_CXX11_ABI=%d' % int(torch._C._GLIBCXX_USE_CXX11_ABI),
           os.path.join(CSRC, 'torch_ext.cpp'), '-o', EXT] + inc + [
        '-L' + tlib, '-lc10', '-lc10_cuda', '-ltorch_cpu', '-ltorch_cuda', '-ltorch', '-ltorch_python',
        '-L' + HERE, '-lmvpnet_b200', '-L/usr/local/cuda/lib64', '-lcudart',
        '-Wl,-rpath,$ORIGIN', '-Wl,-rpath,' + tlib, '-Wl,--no-as-needed']
    _run(cmd, verbose)
    return EXT


def build_all(force=False, verbose=False):
    return build_lib(force, verbose), build_ext(force, verbose)


if __name__ == '__main__':
    force = '--force' in sys.argv
    print(build_all(force=force, verbose=True))
