"""Build libkarman_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

The shared library exports only the C ABI of ``include/karman_b200.h``.  It is compiled with
``-fmad=false`` so that element-wise arithmetic is bit-identical to the (FMA-free) CPU oracle;
the kernels are HBM-bound, so the extra DMUL/DADD issue slots are free.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libkarman_b200.so')
SOURCES = ['kt_api.cu', 'kt_krylov.cu', 'kt_simple.cu', 'kt_channel.cu', 'kt_cluster.cu', 'kt_coop.cu']
HEADERS = ['kt_common.cuh', 'kt_handle.h', os.path.join('..', '..', 'include', 'karman_b200.h')]

NVCC_FLAGS = [
    '-O3', '-std=c++17', '-lineinfo',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-fmad=false',
    '-Xcompiler', '-fPIC', '-shared',
    '-cudart', 'static',
]


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: libkarman_b200.so cannot be built')


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile every .cu into one shared library next to this file.  Returns the path.
    ``out`` + ``extra_flags`` (-D...) build a tuning variant elsewhere (tools/variants.py)."""
    if out is None and not force and not needs_build():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + ['-o', out or LIB] + srcs
    if verbose:
        print(' '.join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libkarman_b200.so')
    if verbose and (res.stdout or res.stderr):
        print(res.stdout + res.stderr)
    return out or LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True,
          extra_flags=(['-Xptxas', '-v'] if '--ptxas' in sys.argv else []))
    print(LIB)
