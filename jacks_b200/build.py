"""Build recipe for libjacks_b200.so (nvcc, sm_100a only).

    python -m jacks_b200.build          # build if stale
    python -m jacks_b200.build --force

The shared library is built IN-TREE (jacks_b200/libjacks_b200.so) so that it travels to the
GPU box with the repository snapshot.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
OBJ = os.path.join(PKG, 'csrc', 'build')
LIB = os.path.join(PKG, 'libjacks_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC,-fvisibility=hidden']
FAST_G = list(range(1, 13))      # kMaxFastG in csrc/em_common.cuh

UNITS = [('capi.cu', 'capi.o', []), ('em_dispatch.cu', 'em_dispatch.o', []), ('em_generic.cu', 'em_generic.o', []), ('em_lines.cu', 'em_lines.o', []),
         ('fp64_probe.cu', 'fp64_probe.o', []), ('kde_pvalues.cu', 'kde_pvalues.o', []),
         ('preprocess_kernels.cu', 'preprocess_kernels.o', []), ('host_io.cpp', 'host_io.o', ['-x', 'cu'])] + \
        [('em_fast_inst.cu', 'em_fast_g%d.o' % g, ['-DJACKS_G=%d' % g]) for g in FAST_G]


def _sources():
    out = [os.path.join(REPO, 'include', 'jacks_b200.h')]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith(('.cu', '.cuh', '.h', '.cpp')):
            out.append(os.path.join(CSRC, f))
    return out


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('build failed: %s\n%s\n%s' % (' '.join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build_library(force=False, verbose=False, variant=None, defines=()):
    """variant/defines: developer A/B builds -- libjacks_b200_<variant>.so compiled with extra -D flags (objects in
    csrc/build/<variant>/); load it with JACKS_B200_LIB=<path> (jacks_b200/_native.py)."""
    global OBJ, LIB
    if variant:
        OBJ = os.path.join(os.environ.get('JACKS_VARIANT_OBJ', '/tmp/jacks_b200_build'), variant)      # out of the tree: the GPU snapshot is size-capped
        LIB = os.path.join(PKG, 'libjacks_b200_%s.so' % variant)
    deps = _sources()
    units = [u for u in UNITS if os.path.exists(os.path.join(CSRC, u[0]))]
    if not force and not _stale(LIB, deps):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [d for d in deps if not d.endswith(('.cu', '.cpp'))]

    def compile_one(u):
        src, obj, extra = os.path.join(CSRC, u[0]), os.path.join(OBJ, u[1]), u[2]
        if force or _stale(obj, [src] + hdrs):
            log = _run([NVCC] + ARCH + FLAGS + ['-D' + d for d in defines] + (['-Xptxas', '-v'] if verbose else []) + extra + ['-c', src, '-o', obj])
            if verbose:
                print(log)
        return obj
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, units))
    _run([NVCC] + ARCH + ['-shared', '-Xcompiler', '-fPIC', '-o', LIB] + objs + ['-cudart', 'static'])
    return LIB


if __name__ == '__main__':
    var = [a.split('=', 1)[1] for a in sys.argv if a.startswith('--variant=')]
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv, variant=var[0] if var else None,
                        defines=[a[2:] for a in sys.argv if a.startswith('-D')]))
