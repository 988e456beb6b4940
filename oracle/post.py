"""NumPy restatement of the reference's display frames and Jacobi sweeps.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows

* ``display.py:17-49``: ``scale_field`` (f_max from -1e9, f_min from 1e9), ``post_process_field``, the (u, v, p, pcor, mdiv) frame;
* ``karman_taichi.py:544-583,662-665``: the monolith's ``post_process_field`` (udiv recomputed in the interior, x3 upsample),
  ``scale_field`` (f_max from 0.0) and the (pcor, udiv, p, u, v) frame;
* ``pressure_valid.py:220-232`` / ``momentum_valid.py:157-170``: one Jacobi sweep, every read taken from the old iterate.

The two frame functions are pinned bit for bit to the reference's own output (``tests/golden/reference_cavity.npz`` and
``reference_channel.npz`` carry the display fields the reference produced).
"""
from __future__ import annotations

import numpy as np


def scale_field(f, fmax0):
    """``display.py:17-26`` (fmax0 = -1e9) / ``karman_taichi.py:573-583`` (fmax0 = 0.0): f_max = the largest of fmax0 and the
    entries, f_min = the smallest of 1e9 and the entries (a NaN never wins ``f[i,j] > f_max``)."""
    f_max = float(np.fmax(np.fmax.reduce(f, axis=None), fmax0))
    f_min = float(np.fmin(np.fmin.reduce(f, axis=None), 1.0e9))
    return (f - f_min) / (f_max - f_min + 1.0e-9)


def class_frame(u, v, p, pcor, mdiv):
    """``Display.post_process_field`` + ``display`` (``display.py:28-49``): returns (img, panels)."""
    nxc, nyc = p.shape
    panels = {'u': 0.5 * (u[0:nxc, :] + u[1:nxc + 1, :]), 'v': 0.5 * (v[:, 0:nyc] + v[:, 1:nyc + 1]), 'p': p.copy(), 'pcor': pcor.copy(),
              'div': mdiv.copy()}
    panels = {k: scale_field(a, -1.0e9) for k, a in panels.items()}
    return np.concatenate([panels[k] for k in ('u', 'v', 'p', 'pcor', 'div')], axis=1), panels


def monolith_frame(u, v, p, pcor, udiv, dx, dy):
    """``post_process_field`` + the GUI frame of ``karman_taichi.py:544-570,662-665``: returns (img, panels, udiv)."""
    nxc, nyc = p.shape
    nx, ny = nxc - 2, nyc - 2
    u_post = 0.5 * (u[0:nxc, :] + u[1:nxc + 1, :])
    v_post = 0.5 * (v[:, 0:nyc] + v[:, 1:nyc + 1])
    udiv = udiv.copy()
    udiv[1:nx + 1, 1:ny + 1] = (u[1:nx + 1, 1:ny + 1] - u[2:nx + 2, 1:ny + 1]) * dy + (v[1:nx + 1, 2:ny + 2] - v[1:nx + 1, 1:ny + 1]) * dx
    up = lambda a: np.repeat(np.repeat(a, 3, axis=0), 3, axis=1)          # disp[i, j] = f[i // 3, j // 3]
    panels = {'u': up(u_post), 'v': up(v_post), 'div': up(udiv), 'p': up(p), 'pcor': up(pcor)}
    panels = {k: scale_field(a, 0.0) for k, a in panels.items()}
    return np.concatenate([panels[k] for k in ('pcor', 'div', 'p', 'u', 'v')], axis=1), panels, udiv


def jacobi_sweep(coef, rhs, x, rows, cols, relax=1.0):
    """One sweep over rows [r0, r1) x cols [c0, c1); returns (x_new, residual)."""
    (r0, r1), (c0, c1) = rows, cols
    I, J = slice(r0, r1), slice(c0, c1)
    acc = -coef[I, J, 1] * x[r0 - 1:r1 - 1, J]
    acc = acc - coef[I, J, 2] * x[r0 + 1:r1 + 1, J]
    acc = acc - coef[I, J, 3] * x[I, c0 + 1:c1 + 1]
    acc = acc - coef[I, J, 4] * x[I, c0 - 1:c1 - 1]
    mid = (acc + rhs[I, J]) / coef[I, J, 0]
    res = float(np.max(np.abs(x[I, J] - mid))) if mid.size else 0.0
    out = x.copy()
    out[I, J] = mid if relax == 1.0 else (1 - relax) * x[I, J] + relax * mid
    return out, res
