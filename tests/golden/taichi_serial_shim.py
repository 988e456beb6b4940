"""A serial, strict-IEEE stand-in for the few Taichi 0.9 calls the reference uses.

TEST INFRASTRUCTURE.  Taichi itself cannot be installed here (no wheel for Python 3.12, no network), so the reference's
Python files cannot run as they are.  Their arithmetic, however, is ordinary Python syntax inside ``@ti.kernel``
functions; this module provides just enough of the ``taichi`` namespace (``field``, ``root.dense/pointer/place``,
``kernel``, ``func``, ``ndrange``, ``grouped``, ``static``, ``sqrt``, ``abs`` ...) for the UNMODIFIED reference sources under
``/root/reference`` to be imported and executed by CPython.  ``make_reference_golden.py`` uses it to produce the golden
vectors under ``tests/golden/reference_*.npz``.

Semantics chosen (and where they differ from a real Taichi run):

* every top-level ``for`` of a kernel runs serially in index order (a legal schedule of Taichi's parallel loops: the
  reference's loops are race-free except for ``sum +=`` reductions, whose order Taichi leaves unspecified -- here it is
  row-major index order);
* arithmetic is IEEE double without FMA contraction or fast-math re-association (``ti.init`` defaults to
  ``fast_math=True`` on LLVM, so a real run may differ in the last bits);
* inside a kernel a field read gives a NumPy scalar, so division by zero yields inf/nan as in compiled Taichi code;
  in Python scope field reads and kernel return values are Python floats, so ``a / 0.0`` raises ``ZeroDivisionError``
  exactly like the reference's host-side divisions (``bicgsolver.py:102,108,118``, ``cgsolver.py:96``);
* an out-of-range read of a dense field gives 0 and an out-of-range write is dropped (Taichi release builds do not
  bounds-check; the reference only ever multiplies such reads by zero coefficients);
* ``ti.root.pointer(...).dense(...).place(A)`` fields are dictionaries keyed by the raw index pair: inactive cells read 0.
"""
from __future__ import annotations

import builtins
import functools
import itertools
import math

import numpy as np

f64 = np.float64
f32 = np.float32
i32 = np.int32
cpu = 'cpu'
gpu = 'gpu'
ij = 'ij'
i = 'i'
j = 'j'

_depth = 0          # > 0 while a kernel body is executing


def init(**kw):
    return None


def template():
    return None


def static(x):
    return x


def data_oriented(cls):
    return cls


def func(fn):
    return fn


def kernel(fn):
    @functools.wraps(fn)
    def run(*a, **k):
        global _depth
        _depth += 1
        try:
            with np.errstate(all='ignore'):
                out = fn(*a, **k)
        finally:
            _depth -= 1
        if isinstance(out, np.generic):
            out = out.item()
        return out
    run.__wrapped_kernel__ = fn
    return run


def sqrt(x):
    if _depth == 0 and not isinstance(x, np.generic):
        return math.sqrt(x)
    with np.errstate(all='ignore'):
        return np.sqrt(np.float64(x))


abs = builtins.abs          # noqa: A001
max = builtins.max          # noqa: A001
min = builtins.min          # noqa: A001


def cast(x, t):
    return t(x)


def ndrange(*args):
    rs = [range(a) if isinstance(a, int) else range(a[0], a[1]) for a in args]
    if len(rs) == 1:
        return iter(rs[0])
    return itertools.product(*rs)


def grouped(f):
    return iter(f._indices(grouped=True))


class Field:
    def __init__(self, dtype, shape=None):
        self.dtype = dtype
        self._a = None
        self._d = None
        self.shape = None
        if shape is not None:
            self._alloc(shape)

    def _alloc(self, shape):
        if isinstance(shape, int):
            shape = (shape,)
        self.shape = tuple(int(s) for s in shape)
        self._a = np.zeros(self.shape, dtype=self.dtype)

    def _alloc_sparse(self, shape):
        self.shape = tuple(int(s) for s in shape)
        self._d = {}

    def _indices(self, grouped=False):
        if self._d is not None:
            raise NotImplementedError('struct-for over a pointer field (the reference has none)')
        if len(self.shape) == 1 and not grouped:
            return range(self.shape[0])
        return np.ndindex(*self.shape)

    def __iter__(self):
        return iter(self._indices())

    def _zero(self):
        return self.dtype(0)

    def __getitem__(self, idx):
        if idx is None:
            v = self._a[()]
        else:
            if not isinstance(idx, tuple):
                idx = (idx,)
            if self._d is not None:
                v = self._d.get(idx)
                if v is None:
                    v = self._zero()
            else:
                for k, n in zip(idx, self.shape):
                    if k < 0 or k >= n:
                        v = self._zero()
                        break
                else:
                    v = self._a[idx]
        return v if _depth else v.item()

    def __setitem__(self, idx, val):
        if idx is None:
            self._a[()] = val
            return
        if not isinstance(idx, tuple):
            idx = (idx,)
        if self._d is not None:
            self._d[idx] = self.dtype(val)
            return
        for k, n in zip(idx, self.shape):
            if k < 0 or k >= n:
                return
        self._a[idx] = val

    def to_numpy(self):
        if self._d is not None:
            raise NotImplementedError
        return self._a.copy()

    def from_numpy(self, a):
        self._a[...] = a

    def fill(self, v):
        self._a[...] = v

    def copy_from(self, other):
        self._a[...] = other._a


def field(dtype, shape=None):
    return Field(dtype, shape)


class _SNode:
    def __init__(self, levels=(), sparse=False):
        self.levels = levels
        self.sparse = sparse

    def _child(self, axes, shape, sparse):
        n = len(axes)
        if isinstance(shape, int):
            shape = (shape,) * n
        assert len(shape) == n
        return _SNode(self.levels + (tuple(shape),), self.sparse or sparse)

    def dense(self, axes, shape):
        return self._child(axes, shape, False)

    def pointer(self, axes, shape):
        return self._child(axes, shape, True)

    def place(self, *fields):
        if not self.levels:
            for f in fields:
                f._a = np.zeros((), dtype=f.dtype)
                f.shape = ()
            return self
        shape = [1] * len(self.levels[0])
        for lv in self.levels:
            assert len(lv) == len(shape)
            shape = [a * b for a, b in zip(shape, lv)]
        for f in fields:
            if self.sparse:
                f._alloc_sparse(shape)
            else:
                f._alloc(shape)
        return self


root = _SNode()


class GUI:
    def __init__(self, *a, **k):
        self.running = True

    def set_image(self, img):
        self.img = img

    def show(self, filename=None):
        return None

    def text(self, *a, **k):
        return None
