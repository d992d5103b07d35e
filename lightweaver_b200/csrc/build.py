"""In-tree build of the CUDA library (and, where the reference headers exist,
the Lightweaver plugin shim).  Explicit nvcc for sm_100a; outputs next to the
package so they travel to the GPU box with the snapshot.

    python -m lightweaver_b200.csrc.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
REF_SRC = os.environ.get('LW_REFERENCE_SRC', '/root/reference/Source')


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_cuda(force=False, verbose=False):
    out = os.path.join(PKG, 'liblwb200.so')
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(('.cu', '.cuh', '.inl'))]
    deps.append(os.path.join(ROOT, 'include', 'lwb200.h'))
    if not force and not _stale(out, deps):
        return out
    cmd = [NVCC, '-O3', '-std=c++17', '-lineinfo', *ARCH, '-Xcompiler', '-fPIC', '-shared',
           '-Xptxas', '-v' if verbose else '-warn-spills',
           '-o', out, os.path.join(HERE, 'lwb200_api.cu'), '-lcudart']
    print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    return out


def build_plugin(force=False):
    """The C++ shim that exports fs_iteration_fns_provider / fs_provider.  It is compiled against the
    reference's HEADERS (the provider structs, Context& and ExtraParams cross that boundary as C++
    types), so it can only be (re)built where a Lightweaver source tree exists (LW_REFERENCE_SRC);
    elsewhere the prebuilt .so is used.  It links NO reference object code -- only liblwb200.so."""
    out = os.path.join(PKG, 'liblwb200_plugin.so')
    src = os.path.join(HERE, 'lwb200_plugin.cpp')
    if not os.path.exists(src):
        return None
    if not os.path.isdir(REF_SRC):
        return out if os.path.exists(out) else None
    deps = [src, os.path.join(ROOT, 'include', 'lwb200.h')]
    if not force and not _stale(out, deps):
        return out
    cmd = ['g++', '-std=c++17', '-O2', '-fPIC', '-shared', '-Wno-sign-compare', '-Wl,--no-undefined',
           '-I', REF_SRC, '-I', os.path.join(ROOT, 'include'), '-o', out, src,
           '-L', PKG, '-llwb200', '-lpthread', '-Wl,-rpath,$ORIGIN']
    print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    return out


def build_all(force=False, verbose=False):
    a = build_cuda(force, verbose)
    b = build_plugin(force)
    return a, b


if __name__ == '__main__':
    print(build_all(force='--force' in sys.argv, verbose='-v' in sys.argv))
