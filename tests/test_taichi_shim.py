"""The serial stand-in for the `taichi` module (tests/golden/taichi_serial_shim.py), exercised on kernels written HERE in
the reference's idiom -- not on the reference's files (executing those is what tests/golden/make_reference_golden.py
does, once approved).  Checks the semantics its docstring promises."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location('taichi_serial_shim', os.path.join(HERE, 'golden', 'taichi_serial_shim.py'))
ti = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ti)


def test_fields_and_placement():
    a = ti.field(ti.f64, shape=(3, 4))
    assert a.shape == (3, 4) and a.to_numpy().sum() == 0.0
    a[1, 2] = 2.5
    assert a[1, 2] == 2.5 and isinstance(a[1, 2], float)          # Python scope: Python floats
    assert a[-1, 0] == 0.0 and a[3, 0] == 0.0                      # out of range reads give 0
    a[7, 7] = 1.0                                                  # ... and out of range writes are dropped
    assert a.to_numpy().sum() == 2.5
    p, q, s = ti.field(ti.f64), ti.field(ti.f64), ti.field(ti.f64)
    ti.root.dense(ti.ij, (2, 5)).place(p, q)
    ti.root.place(s)
    assert p.shape == (2, 5) and q.shape == (2, 5) and s.shape == ()
    s[None] = 3.0
    assert s[None] == 3.0
    A = ti.field(ti.f64)
    ti.root.pointer(ti.ij, (4, 4)).dense(ti.ij, (16, 16)).place(A)
    assert A.shape == (64, 64) and A[5, 6] == 0.0
    A[5, 6] = 1.5
    A[0, -1] = 9.0                                                 # raw keys are kept apart, never wrapped
    assert A[5, 6] == 1.5 and A[0, -1] == 9.0 and A[0, 63] == 0.0
    c = ti.field(ti.i32, shape=(2, 2))
    c[0, 0] = 1.9                                                  # integer fields truncate like Taichi's cast
    assert c[0, 0] == 1 and isinstance(c[0, 0], int)


def test_kernel_semantics():
    x = ti.field(ti.f64, shape=(4, 3))
    y = ti.field(ti.f64, shape=(4, 3))
    x.from_numpy(np.arange(12, dtype=np.float64).reshape(4, 3))

    @ti.data_oriented
    class K:
        @ti.kernel
        def shift(self):
            for i, j in ti.ndrange((1, 4), 3):
                y[i, j] = x[i - 1, j] + 0.5 * x[i, j + 1]            # x[i, 3] is out of range -> 0

        @ti.kernel
        def total(self, f: ti.template()) -> ti.f64:
            s = 0.0
            for I in ti.grouped(f):
                s += f[I]
            return s

        @ti.kernel
        def divide(self) -> ti.f64:
            return 1.0 / x[0, 0]                                   # x[0, 0] == 0: inf inside a kernel, no exception

        @ti.kernel
        def struct_for(self):
            for i, j in y:
                y[i, j] = ti.sqrt(ti.abs(y[i, j]))
    k = K()
    k.shift()
    xn = x.to_numpy()
    want = np.zeros((4, 3))
    want[1:, :] = xn[:-1, :]
    want[1:, :2] += 0.5 * xn[1:, 1:]
    np.testing.assert_array_equal(y.to_numpy(), want)
    assert k.total(x) == 66.0 and isinstance(k.total(x), float)    # kernel results come back as Python floats
    assert k.divide() == float('inf')
    with pytest.raises(ZeroDivisionError):                          # ... but Python-scope division raises, as in the reference
        _ = 1.0 / x[0, 0]
    k.struct_for()
    np.testing.assert_array_equal(y.to_numpy(), np.sqrt(np.abs(want)))
    assert ti.sqrt(4.0) == 2.0 and list(ti.static([1, 2])) == [1, 2]


def test_serial_sum_order_matches_oracle_serial_mode():
    from oracle import krylov
    rng = np.random.default_rng(0)
    a = rng.standard_normal((7, 9))
    f = ti.field(ti.f64, shape=a.shape)
    f.from_numpy(a)

    @ti.kernel
    def reduce(p: ti.template(), q: ti.template()) -> ti.f64:
        s = 0.0
        for I in ti.grouped(p):
            s += p[I] * q[I]
        return s
    krylov.set_summation('serial')
    try:
        assert reduce(f, f) == krylov.reduce(a, a)                  # bit for bit
    finally:
        krylov.set_summation('pairwise')
