"""oracle/device_order.py on the CPU: the plan functions are host-only, so the restated summation orders can be checked
for coverage without a GPU -- with integer-valued products every summation order is exact, so each plan must return the
plain integer sum (every owned element visited exactly once, nothing else).  That the ORDER equals the device's is what the
-m gpu twin (tests/test_gpu_exact.py) proves, bit for bit."""
import numpy as np
import pytest

from oracle import krylov
from oracle.device_order import DeviceOrder, SlabOrder, tree, seq

SHAPES = [(5, 7), (16, 16), (37, 53), (52, 52), (130, 130), (322, 66), (66, 323), (642, 130), (258, 1030), (1026, 514)]


def test_tree_and_seq_shapes():
    a = np.arange(32.0)
    assert tree(a) == a.sum() and seq(a[None, :])[0] == a.sum()
    # the tree really is the shuffle tree: ((a0 + a2) + (a1 + a3)) for four leaves
    b = np.array([1e16, 1.0, -1e16, 1.0])
    assert tree(b) == (1e16 + -1e16) + (1.0 + 1.0)
    assert seq(b[None, :])[0] == ((1e16 + 1.0) + -1e16) + 1.0


@pytest.mark.parametrize('shape', SHAPES)
@pytest.mark.parametrize('enable', [0, 1])
def test_every_plan_visits_each_element_once(shape, enable, lib_path):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    prod = rng.integers(-1000, 1000, size=shape).astype(np.float64)
    exact = float(prod.sum())
    for is_cg, pre, warm in ((False, False, False), (True, False, False), (False, True, True)):
        order = DeviceOrder(shape, is_cg=is_cg, pre=pre, warm=warm, enable=enable)
        assert enable == 1 or order.path == 0
        kinds = ('cg_init', 'cg_alpha', 'cg_end') if is_cg else ('init', 'alpha', 'omega', 'end', 'resid', 'dot')
        for kind in kinds:
            assert order(kind, prod) == exact, (shape, enable, order.path, kind)


def test_paths_cover_the_three_solvers(lib_path):
    paths = {shape: DeviceOrder(shape).path for shape in SHAPES}
    assert paths[(16, 16)] == 1 and paths[(130, 130)] == 1      # cluster, registers
    assert paths[(322, 66)] == 2                                # cluster, shared memory
    assert paths[(642, 130)] == 3                               # cooperative
    assert paths[(1026, 514)] == 0                              # streaming


@pytest.mark.parametrize('nranks', [2, 3, 5])
def test_slab_order_covers_owned_rows(nranks, lib_path):
    shape = (67, 41)
    base, rem = divmod(shape[0], nranks)
    bounds = [0]
    for r in range(nranks):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    prod = np.random.default_rng(nranks).integers(-500, 500, size=shape).astype(np.float64)
    order = SlabOrder(shape, bounds)
    for kind in ('init', 'alpha', 'omega', 'end'):
        assert order(kind, prod) == float(prod.sum())


def test_device_summation_mode_runs_the_oracle_solvers(lib_path):
    """Different orders, same mathematics: the solves agree to 1e-8; CG stops within +-1 iteration, BiCGSTAB within a few
    (its stopping iteration is sensitive to the last bits of the inner products: 53 vs 56 here between NumPy's pairwise
    and the device's order -- the reason the GPU parity test pins the ORDER and then demands equality, not a band)."""
    coef, b = krylov.poisson_sin_system(30, 0.1)
    x0, it0, st0 = krylov.bicgstab_solve(coef, b, eps=1e-10)
    c0, ic0, _ = krylov.cg_solve(coef, b, eps=1e-10)
    try:
        krylov.set_summation('device', DeviceOrder(b.shape))
        x1, it1, st1 = krylov.bicgstab_solve(coef, b, eps=1e-10)
        krylov.set_summation('device', DeviceOrder(b.shape, is_cg=True))
        c1, ic1, _ = krylov.cg_solve(coef, b, eps=1e-10)
    finally:
        krylov.set_summation('pairwise')
    assert st0 == st1 == 0 and abs(it0 - it1) <= 5 and abs(ic0 - ic1) <= 1
    assert np.linalg.norm(x1 - x0) / np.linalg.norm(x0) < 1e-8
    assert np.linalg.norm(c1 - c0) / np.linalg.norm(c0) < 1e-8
