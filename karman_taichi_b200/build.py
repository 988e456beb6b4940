"""Build libkarman_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

The shared library exports only the C ABI of ``include/karman_b200.h``.  It is compiled with
``-fmad=false`` so that element-wise arithmetic is bit-identical to the (FMA-free) CPU oracle;
the kernels are HBM-bound, so the extra DMUL/DADD issue slots are free.

Every ``csrc/*.cu`` is compiled to its own object under ``karman_taichi_b200/_obj/`` (in parallel, and only when
the source or a header changed), then linked.  Objects are build products: git-ignored (``*.o``) and listed in
``.gpurunignore`` -- only the linked ``.so`` travels to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, '_obj')
LIB = os.path.join(HERE, 'libkarman_b200.so')
SOURCES = ['kt_api.cu', 'kt_krylov.cu', 'kt_simple.cu', 'kt_channel.cu', 'kt_cluster.cu', 'kt_coop.cu', 'kt_post.cu']
HEADERS = ['kt_common.cuh', 'kt_handle.h', os.path.join('..', '..', 'include', 'karman_b200.h')]

COMPILE_FLAGS = [
    '-O3', '-std=c++17', '-lineinfo',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-fmad=false',
    '-Xcompiler', '-fPIC',
]
LINK_FLAGS = ['-shared', '-cudart', 'static', '-gencode', 'arch=compute_100a,code=sm_100a']
NVCC_FLAGS = COMPILE_FLAGS + LINK_FLAGS      # (kept: what one monolithic nvcc call would take)


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: libkarman_b200.so cannot be built')


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _header_mtime():
    return max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS if os.path.exists(os.path.join(CSRC, h)))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in _sources() + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd, verbose):
    if verbose:
        print(' '.join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libkarman_b200.so: ' + ' '.join(cmd))
    if verbose and (res.stdout or res.stderr):
        print(res.stdout + res.stderr)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile every .cu and link them into one shared library next to this file.  Returns the path.
    ``out`` + ``extra_flags`` (-D...) build a tuning variant elsewhere (tools/variants.py)."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    extra = list(extra_flags)
    # objects of a variant build live apart from the shipped library's
    tag = hashlib.sha1(' '.join(extra).encode()).hexdigest()[:8] if extra else 'default'
    odir = os.path.join(OBJ, tag)
    os.makedirs(odir, exist_ok=True)
    newest_dep = max(_header_mtime(), os.path.getmtime(os.path.abspath(__file__)))
    jobs, objs = [], []
    for s in _sources():
        src, obj = os.path.join(CSRC, s), os.path.join(odir, s[:-3] + '.o')
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_dep)
        if stale:
            jobs.append([nvcc] + COMPILE_FLAGS + extra + ['-c', src, '-o', obj])
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            list(ex.map(lambda c: _run(c, verbose), jobs))
    _run([nvcc] + LINK_FLAGS + ['-Xcompiler', '-fPIC', '-o', out or LIB] + objs, verbose)
    return out or LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True,
          extra_flags=(['-Xptxas', '-v'] if '--ptxas' in sys.argv else []))
    print(LIB)
