"""Build the reference's own CUDA extensions for sm_100a into oracle/_ref/ (git-ignored, travels to
the GPU box).  TEST INFRASTRUCTURE: gives the GPU tests the reference's real kernels to compare
against (tests/test_gpu_vs_reference_kernels.py) and lets bench.py time "the reference algorithm on
B200".  The sources are compiled where they lie under /root/reference/mvpnet/ops/cuda — nothing is
copied into the repo — with one injected include directory (oracle/ref_compat: a <THC/THC.h> shim,
that header no longer exists in PyTorch 2.x) and a forced include of the same shim for the files that
use the glog-style CHECK_* macros without including THC.

Recipe per module (mirrors mvpnet/ops/setup.py:13-67, which only lists sources + `nvcc -O2`):
    nvcc -O2 -gencode arch=compute_100a,code=sm_100a -c <name>_kernel.cu
    g++  -O2 -c <name>.cpp ; g++ -shared -> oracle/_ref/<name>_cuda*.so
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/mvpnet/ops/cuda'
OUT = os.path.join(HERE, '_ref')
MODULES = {'fps_cuda': 'fps', 'ball_query_cuda': 'ball_query', 'ball_query_distance_cuda': 'ball_query_distance',
           'group_points_cuda': 'group_points', 'knn_distance_cuda': 'knn_distance', 'interpolate_cuda': 'interpolate'}


def main():
    if not os.path.isdir(REF):
        print('build_ref: %s not present, nothing to do' % REF)
        return 0
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    tlib = os.path.join(os.path.dirname(torch.__file__), 'lib')
    inc = ['-I' + os.path.join(HERE, 'ref_compat')]
    for p in ce.include_paths() + [sysconfig.get_paths()['include'], '/usr/local/cuda/include']:
        inc += ['-isystem', p]
    shim = os.path.join(HERE, 'ref_compat', 'THC', 'THC.h')
    abi = '-D_GLIBCXX_USE_CXX11_ABI=%d' % int(torch._C._GLIBCXX_USE_CXX11_ABI)
    suffix = sysconfig.get_config_var('EXT_SUFFIX') or '.so'
    failed = []
    for mod, stem in MODULES.items():
        target = os.path.join(OUT, mod + suffix)
        srcs = [os.path.join(REF, stem + '.cpp'), os.path.join(REF, stem + '_kernel.cu')]
        if os.path.exists(target) and all(os.path.getmtime(target) > os.path.getmtime(s) for s in srcs + [shim]):
            continue
        defs = ['-DTORCH_EXTENSION_NAME=' + mod, '-DTORCH_API_INCLUDE_EXTENSION_H', abi]
        o_cu, o_cpp = os.path.join(OUT, stem + '_kernel.o'), os.path.join(OUT, stem + '.o')
        cmds = [
            ['nvcc', '-O2', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
             '-include', shim, '-c', srcs[1], '-o', o_cu] + defs + inc,
            ['/usr/bin/g++', '-O2', '-std=c++17', '-fPIC', '-include', shim, '-c', srcs[0], '-o', o_cpp] + defs + inc,
            ['/usr/bin/g++', '-shared', o_cu, o_cpp, '-o', target, '-L' + tlib, '-lc10', '-lc10_cuda', '-ltorch_cpu', '-ltorch_cuda',
             '-ltorch', '-ltorch_python', '-L/usr/local/cuda/lib64', '-lcudart', '-Wl,-rpath,' + tlib],
        ]
        for cmd in cmds:
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write('build_ref: %s failed:\n%s\n' % (mod, (r.stdout + r.stderr)[-3000:]))
                failed.append(mod)
                break
        for o in (o_cu, o_cpp):
            if os.path.exists(o):
                os.remove(o)
    print('build_ref: built %d/%d reference extension modules into %s%s' %
          (len(MODULES) - len(failed), len(MODULES), OUT, (' (failed: %s)' % failed) if failed else ''))
    return 1 if failed else 0


if __name__ == '__main__':
    sys.exit(main())
