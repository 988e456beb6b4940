"""NumPy restatement of the reference's matrix-free CG / BiCGSTAB solvers.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows, line by line,

* ``cgsolver.py:31-114``   (``CGSolver``: init, reduce, compute_Ap, update_*, solve)
* ``bicgsolver.py:21-135`` (``BICGSolver``: unpreconditioned BiCGSTAB, x0 = 0)
* ``karman_taichi.py:342-456`` (monolith ``bicgstab``: Jacobi right-preconditioned,
  warm-started) -- in 5-point stencil form, see ``jacobi_bicgstab``.

Array convention (``cgsolver.py:56-60``): ``coef[i, j, k]`` with
k = 0 ap (centre), 1 aw (i-1), 2 ae (i+1), 3 an (j+1), 4 as (j-1); vectors are
``(NX, NY)`` float arrays including the zero-coefficient ghost ring.  A neighbour that
falls outside the array contributes 0 (the reference multiplies whatever it reads
there by a ghost-row coefficient that is never written and therefore 0).

The stencil sum is evaluated left to right without FMA contraction, exactly in the
order the reference writes it, so element-wise results are bit-identical to a
``-fmad=false`` device kernel.  Dot products use ``np.sum`` (pairwise); the reference's
own summation order is thread-nondeterministic, so dots agree to rounding only.
"""
from __future__ import annotations

import math

import numpy as np

# status codes shared with include/karman_b200.h (KT_STATUS_*)
CONVERGED = 0
MAX_ITER = 1
RHO_ZERO = 2        # bicgsolver.py:96-98 '>>> BICG failed at first place...'
ALPHA_DIV_ZERO = 3  # bicgsolver.py:108 unguarded rho / 0.0
OMEGA_DIV_ZERO = 4  # bicgsolver.py:118 unguarded upper / 0.0
NON_FINITE = 5
OMEGA_ZERO = 6      # karman_taichi.py:451-454 'Omega = 0.0'


def _shift(x, di, dj):
    """x[i+di, j+dj] with zeros where the index leaves the array."""
    out = np.zeros_like(x)
    nx, ny = x.shape
    si = slice(max(0, -di), nx - max(0, di))
    ti = slice(max(0, di), nx - max(0, -di))
    sj = slice(max(0, -dj), ny - max(0, dj))
    tj = slice(max(0, dj), ny - max(0, -dj))
    out[si, sj] = x[ti, tj]
    return out


def spmv(coef, x):
    """y = A x, ``cgsolver.py:56-60`` / ``bicgsolver.py:65-76`` (same expression)."""
    y = coef[:, :, 0] * x
    y = y + coef[:, :, 1] * _shift(x, -1, 0)
    y = y + coef[:, :, 2] * _shift(x, +1, 0)
    y = y + coef[:, :, 3] * _shift(x, 0, +1)
    y = y + coef[:, :, 4] * _shift(x, 0, -1)
    return y


def spmv_monolith(coef, x):
    """y = A x in the monolith's term order (``karman_taichi.py:362-367,401-406,421-426``): its product walks the
    flat offsets ``[i-ny-1, i-ny, i-1, i, i+1, i+ny, i+ny+1]`` of a row and accumulates from 0.0, i.e. in stencil terms
    (i-1) W, (j-1) S, centre, (j+1) N, (i+1) E -- the two outer offsets never hold a non-zero.  Same five
    products as ``spmv`` but a different rounding sequence: the reference-run fixtures
    (``tests/golden/reference_channel.npz``) pin this order bit for bit."""
    y = 0.0 + coef[:, :, 1] * _shift(x, -1, 0)            # 'Ap[i] = 0.0' then '+=' (:402-406)
    y = y + coef[:, :, 4] * _shift(x, 0, -1)
    y = y + coef[:, :, 0] * x
    y = y + coef[:, :, 3] * _shift(x, 0, +1)
    y = y + coef[:, :, 2] * _shift(x, +1, 0)
    return y


_SERIAL = False
_DEVICE = None


def set_summation(mode, device_order=None):
    """'pairwise' (default): ``np.sum``.  'serial': strictly sequential in row-major index order, starting from 0.0 --
    the order a serial execution of the reference's ``sum += p[I] * q[I]`` loop (``cgsolver.py:51-53``) or of the
    monolith's scalar accumulations (``karman_taichi.py:376-378,384-385``) takes; bit-identical to the C port run with
    one thread (``tests/test_oracle_cport.py``) and to the reference sources executed under the serial Taichi stand-in
    (``tests/golden/reference_*.npz``).  Real Taichi leaves the order of that reduction unspecified.
    'device': ``device_order(kind, products)`` (``oracle/device_order.py``) -- the order libkarman_b200 adds in; ``kind``
    names the inner product ('init', 'alpha', 'omega', 'end', 'resid', 'cg_init', 'cg_alpha', 'cg_end', 'dot')."""
    global _SERIAL, _DEVICE
    if mode not in ('pairwise', 'serial', 'device'):
        raise ValueError(mode)
    if mode == 'device' and device_order is None:
        raise ValueError("set_summation('device', ...) needs a DeviceOrder")
    _SERIAL = mode == 'serial'
    _DEVICE = device_order if mode == 'device' else None


def _sum(a, kind='dot'):
    if _DEVICE is not None:
        return np.float64(_DEVICE(kind, a))
    if _SERIAL:
        flat = np.ascontiguousarray(a, dtype=np.float64).ravel()
        return np.float64(np.add.accumulate(flat)[-1]) if flat.size else np.float64(0.0)   # ((0 + a0) + a1) + ...
    return np.sum(a)


def reduce(p, q, kind='dot'):
    """``cgsolver.py:49-54``: sum over the whole array, ghosts included."""
    return float(_sum(p * q, kind))


def _div(a, b):
    """Python-float division as the reference performs it on host scalars."""
    return a / b  # raises ZeroDivisionError exactly like the reference


def cg_solve(coef, b, eps=1e-8, max_iter=None, history=None):
    """``CGSolver.solve`` (``cgsolver.py:83-114``).  Returns (x, iters, status).

    ``iters`` counts executed loop bodies (the reference's ``i + 1`` at the break).
    ``history`` (a list) receives the same strings the reference appends
    (``cgsolver.py:88,111``).
    """
    nx, ny = b.shape
    steps = nx * ny if max_iter is None else max_iter  # cgsolver.py:14
    # init(): cgsolver.py:31-39
    r = b.copy()
    x = np.zeros_like(b)
    p = np.zeros_like(b)
    initial_rTr = reduce(r, r, 'cg_init')
    if history is not None:
        history.append(f'{math.sqrt(initial_rTr):e}\n')
    old_rTr = initial_rTr
    beta = 0.0  # ti.field zero-initialised; first update_p uses beta = 0 (cgsolver.py:90)
    p = r + beta * p
    iters = 0
    status = MAX_ITER
    for i in range(steps):
        iters = i + 1
        Ap = spmv(coef, p)                 # :94
        pAp = reduce(p, Ap, 'cg_alpha')    # :95
        alpha = _div(old_rTr, pAp)         # :96
        x = x + alpha * p                  # :98 (update_x :68-71)
        r = r - alpha * Ap                 # :99 (update_r :73-76)
        new_rTr = reduce(r, r, 'cg_end')   # :101
        if math.sqrt(new_rTr) < math.sqrt(initial_rTr) * eps:  # :102
            status = CONVERGED
            break
        beta = _div(new_rTr, old_rTr)      # :107
        p = r + beta * p                   # :109 (update_p :78-81)
        old_rTr = new_rTr
        if history is not None:
            history.append(f'{math.sqrt(new_rTr):e}\n')
    return x, iters, status


def bicgstab_solve(coef, b, eps=1e-8, max_iter=None, history=None):
    """``BICGSolver.solve`` (``bicgsolver.py:88-135``).  Returns (x, iters, status).

    Unpreconditioned (``p_hat = p``, ``s_hat = s``: ``bicgsolver.py:48-58``), x0 = 0,
    r_tld = b (``:21-36``).  ``iters`` = number of loop bodies entered and completed up
    to and including the converging one; the ``rho == 0`` exit (``:96-98``) does not
    count its aborted body.  A zero ``alpha_lower`` / ``omega_lower`` raises
    ``ZeroDivisionError`` like the reference's Python division (``:108,118``).
    """
    nx, ny = b.shape
    steps = nx * ny if max_iter is None else max_iter
    r = b.copy()
    r_tld = b.copy()
    x = np.zeros_like(b)
    p = np.zeros_like(b)
    Ap = np.zeros_like(b)
    omega = alpha = beta = rho_1 = 1.0
    initial_rTr = reduce(r, r, 'init')                       # :90
    iters = 0
    status = MAX_ITER
    for i in range(steps):
        rho = reduce(r, r_tld, 'init' if i == 0 else 'end')  # :95 (the device forms it in the pass that produced r)
        if rho == 0.0:                                       # :96-98
            status = RHO_ZERO
            break
        if i == 0:
            p = r.copy()                                     # :100
        else:
            beta = _div(rho, rho_1) * _div(alpha, omega)     # :102
            p = r + beta * (p - omega * Ap)                  # :103 (update_p :43-46)
        Ap = spmv(coef, p)                                   # :106
        alpha_lower = reduce(r_tld, Ap, 'alpha')             # :107
        alpha = _div(rho, alpha_lower)                       # :108
        s = r - alpha * Ap                                   # :110 (update_s :60-63)
        t = spmv(coef, s)                                    # :113-114
        omega_upper = reduce(t, s, 'omega')                  # :116
        omega_lower = reduce(t, t, 'omega')                  # :117
        omega = _div(omega_upper, omega_lower)               # :118
        x = x + (alpha * p + omega * s)                      # :120 (update_x :78-81)
        r = s - omega * t                                    # :121 (update_r :83-86)
        rTr = reduce(r, r, 'end')                            # :123
        iters = i + 1
        if history is not None:
            history.append(f'{math.sqrt(rTr):e}\n')
        if math.sqrt(_div(rTr, initial_rTr)) < eps:          # :130
            status = CONVERGED
            break
        rho_1 = rho                                          # :135
    return x, iters, status


def jacobi_bicgstab(coef, M, b, x0, eps=1e-8, max_iter=None, n=None):
    """Monolith ``bicgstab`` kernel (``karman_taichi.py:342-456``) in stencil form.

    Right-preconditioned with ``p_hat = p / M``, ``s_hat = s / M`` (``:399,419``; M is the
    separately stored ``M[k,k]`` plane, which is *not* always diag(A): SURVEY A.4),
    warm start ``r = b - A x0`` (``:362-371``), at most ``100 * n`` iterations (``:382``),
    exits on ``rho == 0`` (``:386-389``), convergence (``:446-449``) or ``omega == 0``
    (``:451-454``).  The 7-offset product of ``:365`` equals the 5-point stencil because
    the two extra offsets never hold a non-zero (SURVEY A.4 [probe], re-checked by
    ``tests/test_oracle_karman.py`` against the literal matrix oracle); its terms are added
    in the monolith's flat-index order (``spmv_monolith``).
    ``n`` is the unknown count of the system (interior of the stencil arrays).
    Division by zero inside the Taichi kernel yields inf/nan rather than raising;
    NumPy float64 scalars reproduce that.
    Returns (x, iters, status); ``iters`` = completed loop bodies.
    """
    one = np.float64(1.0)
    if n is None:
        n = int(np.count_nonzero(M))
    steps = 100 * n if max_iter is None else max_iter
    x = x0.copy()
    Minv_safe = np.where(M != 0.0, M, one)  # ghost ring: M = 0 there, vectors stay 0
    Ax = spmv_monolith(coef, x)
    r = b - Ax
    r_tld = r.copy()
    residual_init = np.float64(_sum(r * r, 'resid'))
    omega = alpha = beta = rho_1 = one
    p = np.zeros_like(b)
    Ap = np.zeros_like(b)
    iters = 0
    status = MAX_ITER
    with np.errstate(all='ignore'):
        for step in range(steps):
            rho = np.float64(_sum(r * r_tld, 'resid' if step == 0 else 'end'))
            if rho == 0.0:
                status = RHO_ZERO
                break
            if step == 0:
                p = r.copy()
            else:
                beta = (rho / rho_1) * (alpha / omega)
                p = r + beta * (p - omega * Ap)
            p_hat = np.where(M != 0.0, 1 / Minv_safe * p, 0.0)   # :399 '1/M[i,i] * p[i]'
            Ap = spmv_monolith(coef, p_hat)
            alpha_lower = np.float64(_sum(r_tld * Ap, 'alpha'))
            alpha = rho / alpha_lower
            s = r - alpha * Ap
            s_hat = np.where(M != 0.0, 1 / Minv_safe * s, 0.0)   # :419
            t = spmv_monolith(coef, s_hat)
            omega_upper = np.float64(_sum(t * s, 'omega'))
            omega_lower = np.float64(_sum(t * t, 'omega'))
            omega = omega_upper / omega_lower
            x = x + (alpha * p_hat + omega * s_hat)              # :437
            r = s - omega * t                                    # :440
            residual = np.float64(_sum(r * r, 'end'))
            iters = step + 1
            if np.sqrt(residual / residual_init) < eps:          # :446 (False on NaN, like Taichi)
                status = CONVERGED
                break
            if omega == 0.0:
                status = OMEGA_ZERO
                break
            if not np.isfinite(residual):
                # deviation: the reference would keep iterating on NaNs up to 100*n times
                status = NON_FINITE
                break
            rho_1 = rho
    return x, iters, status


# ---------------------------------------------------------------------------------------------
# Synthetic systems of SURVEY.md section 8(d)
# ---------------------------------------------------------------------------------------------

def poisson_sin_system(n, offset=0.1, dtype=np.float64):
    """S1 'Poisson-sin' after ``benchmark/cg_to_sparse/cgsolver.py:33-56``.

    Shape (n+2, n+2) with a zero-coefficient ghost ring; interior ap = 4 + offset,
    off-diagonals -1 (dropped toward the ring, i.e. Dirichlet), b = sin(2 pi x) sin(2 pi y)
    with x = (i-1)/n (the reference's ``(i - N_ext) * 2 / N_tot``).
    """
    coef = np.zeros((n + 2, n + 2, 5), dtype=dtype)
    b = np.zeros((n + 2, n + 2), dtype=dtype)
    coef[1:n + 1, 1:n + 1, 0] = 4.0 + offset
    coef[1:n + 1, 1:n + 1, 1:5] = -1.0
    coef[1, :, 1] = 0.0
    coef[n, :, 2] = 0.0
    coef[:, n, 3] = 0.0
    coef[:, 1, 4] = 0.0
    xl = np.arange(n, dtype=np.float64) * 2.0 / (2 * n)
    s = np.sin(2.0 * np.pi * xl)
    b[1:n + 1, 1:n + 1] = np.outer(s, s).astype(dtype)
    return coef, b


def random_dd_system(nx, ny, seed=0, dtype=np.float64):
    """S2 random diagonally dominant non-symmetric 5-point system after
    ``debug/ti_bicg_test.py:21-41`` (off-diagonals -U(0,1), diag = -sum(offdiag) + 0.1,
    b = U(0,1)); seeded, with ghost ring, shape (nx+2, ny+2)."""
    rng = np.random.default_rng(seed)
    coef = np.zeros((nx + 2, ny + 2, 5), dtype=np.float64)
    b = np.zeros((nx + 2, ny + 2), dtype=np.float64)
    coef[1:nx + 1, 1:ny + 1, 1:5] = -rng.random((nx, ny, 4))
    coef[1, :, 1] = 0.0
    coef[nx, :, 2] = 0.0
    coef[:, ny, 3] = 0.0
    coef[:, 1, 4] = 0.0
    coef[1:nx + 1, 1:ny + 1, 0] = -coef[1:nx + 1, 1:ny + 1, 1:5].sum(axis=2) + 0.1
    b[1:nx + 1, 1:ny + 1] = rng.random((nx, ny))
    return coef.astype(dtype), b.astype(dtype)


def to_csr(coef):
    """Assemble the stencil arrays into a SciPy CSR matrix over all NX*NY entries
    (row-major k = i*NY + j).  Used only to cross-check the oracle against SciPy."""
    import scipy.sparse as sp
    nx, ny, _ = coef.shape
    n = nx * ny
    idx = np.arange(n).reshape(nx, ny)
    rows, cols, vals = [], [], []

    def add(k, di, dj):
        si = slice(max(0, -di), nx - max(0, di))
        sj = slice(max(0, -dj), ny - max(0, dj))
        ti = slice(max(0, di), nx - max(0, -di))
        tj = slice(max(0, dj), ny - max(0, -dj))
        rows.append(idx[si, sj].ravel())
        cols.append(idx[ti, tj].ravel())
        vals.append(coef[si, sj, k].ravel())

    add(0, 0, 0)
    add(1, -1, 0)
    add(2, +1, 0)
    add(3, 0, +1)
    add(4, 0, -1)
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    return A
