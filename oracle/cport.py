"""ctypes wrapper of oracle/c/libkt_oracle.so (the C/OpenMP port of the reference's Krylov solvers).

TEST / BASELINE INFRASTRUCTURE (see ``oracle/__init__.py``).  ``bench.py`` times this as the CPU baseline
("port": the reference's Taichi ``ti.cpu`` path itself cannot run offline), tests cross-check it against
the NumPy oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CDIR = os.path.join(HERE, 'c')
LIB = os.path.join(CDIR, 'libkt_oracle.so')

_lib = None


def build():
    src = os.path.join(CDIR, 'kt_oracle.c')
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(['make', '-C', CDIR], check=True, capture_output=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        dp = np.ctypeslib.ndpointer(np.float64, flags='C_CONTIGUOUS')
        L.ko_reduce.restype = C.c_double
        L.ko_reduce.argtypes = [dp, dp, C.c_long]
        L.ko_set_threads.restype = None
        L.ko_set_threads.argtypes = [C.c_int]
        L.ko_spmv.restype = None
        L.ko_spmv.argtypes = [dp, dp, dp, C.c_int, C.c_int]
        for fn in (L.ko_cg, L.ko_bicgstab):
            fn.restype = C.c_int
            fn.argtypes = [dp, dp, dp, C.c_int, C.c_int, C.c_double, C.c_long, dp, C.POINTER(C.c_long), C.POINTER(C.c_double)]
        _lib = L
    return _lib


def set_threads(n):
    """Thread count of every later call (1: strictly sequential sums in index order)."""
    os.environ['OMP_NUM_THREADS'] = str(n)
    load().ko_set_threads(int(n))


def spmv(coef, x):
    L = load()
    y = np.empty_like(x)
    L.ko_spmv(np.ascontiguousarray(coef), np.ascontiguousarray(x), y, x.shape[0], x.shape[1])
    return y


def reduce(p, q):
    return load().ko_reduce(np.ascontiguousarray(p).ravel(), np.ascontiguousarray(q).ravel(), p.size)


def _solve(fn, planes, coef, b, eps, max_iter, work=None):
    nx, ny = b.shape
    x = np.zeros_like(b)
    if work is None:
        work = np.zeros(planes * nx * ny)
    it, rel = C.c_long(), C.c_double()
    st = fn(np.ascontiguousarray(coef), np.ascontiguousarray(b), x, nx, ny, float(eps),
            nx * ny if max_iter is None else int(max_iter), work, C.byref(it), C.byref(rel))
    return x, it.value, st, rel.value


def cg_solve(coef, b, eps=1e-8, max_iter=None, work=None):
    return _solve(load().ko_cg, 3, coef, b, eps, max_iter, work)


def bicgstab_solve(coef, b, eps=1e-8, max_iter=None, work=None):
    return _solve(load().ko_bicgstab, 9, coef, b, eps, max_iter, work)
