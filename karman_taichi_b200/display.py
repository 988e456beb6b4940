"""``Display`` -- the reference's post-processing / image dump of the solver fields on the device.

Two frames exist in the reference, both produced here by ``kt_display_frame`` (``csrc/kt_post.cu``):

* the class version (``display.py:4-49``, ``Display(SIMPLESolver)``): cell-centre interpolation ``0.5 (u[i,j] + u[i+1,j])``,
  ``0.5 (v[i,j] + v[i,j+1])``, the five panels (u, v, p, pcor, mdiv) each scaled ``(f - min) / (max - min + 1e-9)`` with
  ``f_max`` starting at -1e9, concatenated along j: an ``(nx+2) x 5 (ny+2)`` image;
* the monolith's (``karman_taichi.py:544-583,662-665``, ``Display(KarmanSolver)``): ``udiv`` recomputed in the interior,
  every panel upsampled x3, ``f_max`` starting at 0.0, panel order (pcor, udiv, p, u, v): ``3 (nx+2) x 15 (ny+2)``.

``ti.GUI`` shows ``img[i, j]`` with i to the right and j upwards; the PNG written here (zlib, 8-bit grey: there is no
ti.GUI / matplotlib in this image) reproduces that orientation.
"""
from __future__ import annotations

import ctypes as C
import struct
import zlib

import numpy as np
import torch

from . import _lib


def write_png_gray(filename, img01):
    """8-bit grayscale PNG of a (height, width) array in [0, 1]."""
    a = np.clip(np.nan_to_num(img01), 0.0, 1.0)
    a = (a * 255.0 + 0.5).astype(np.uint8)
    h, w = a.shape
    raw = b''.join(b'\x00' + a[r].tobytes() for r in range(h))

    def chunk(tag, data):
        c = struct.pack('>I', len(data)) + tag + data
        return c + struct.pack('>I', zlib.crc32(tag + data) & 0xffffffff)

    with open(filename, 'wb') as f:
        f.write(b'\x89PNG\r\n\x1a\n')
        f.write(chunk(b'IHDR', struct.pack('>IIBBBBB', w, h, 8, 0, 0, 0, 0)))
        f.write(chunk(b'IDAT', zlib.compress(raw, 6)))
        f.write(chunk(b'IEND', b''))


class Display:
    def __init__(self, solver, *args):
        self.solver = solver
        self.nx, self.ny = solver.nx, solver.ny
        self.real = solver.real
        self.monolith = not hasattr(solver, 'mdiv')          # KarmanSolver carries udiv, SIMPLESolver mdiv
        self.scale = 3 if self.monolith else 1
        self.minmax = None                                   # {name: (min, max)} of the last frame
        self._img = None
        self.udisp = self.vdisp = self.pdisp = self.pcordisp = self.mdivdisp = None

    def post_process_field(self):
        """``display.py:28-40`` / ``karman_taichi.py:544-570``: fills the five display panels (device -> NumPy views of one
        image) and returns the concatenated image ``img[i, j]``."""
        s = self.solver
        S, nxc, nyc = self.scale, self.nx + 2, self.ny + 2
        div = s.udiv if self.monolith else s.mdiv
        if self._img is None:
            self._img = torch.empty((S * nxc, 5 * S * nyc), dtype=s.p.dtype, device=s.p.device)
        P = _lib.FrameParams(self.nx, self.ny, s.u.pitch, s.v.pitch, s.p.pitch, S, 1 if self.monolith else 0, float(s.dx), float(s.dy))
        mm = (C.c_double * 10)()
        ptr = lambda f: C.c_void_p(f.ptr)
        _lib.check(s._lib.kt_display_frame(s._h, C.byref(P), ptr(s.u), ptr(s.v), ptr(s.p), ptr(s.pcor), ptr(div),
                                           C.c_void_p(self._img.data_ptr()), mm))
        names = ('u', 'v', 'p', 'pcor', 'div')
        self.minmax = {n: (mm[5 + k], mm[k]) for k, n in enumerate(names)}
        img = self._img.cpu().numpy()
        w = S * nyc
        order = ('pcor', 'div', 'p', 'u', 'v') if self.monolith else names
        panels = {n: img[:, k * w:(k + 1) * w] for k, n in enumerate(order)}
        self.udisp, self.vdisp, self.pdisp, self.pcordisp, self.mdivdisp = (panels[n] for n in names)
        return img

    def image(self):
        return self.post_process_field()

    def display(self, filename):
        """``display.py:42-49`` / ``karman_taichi.py:662-676``: the frame saved to ``filename``; returns ``img[i, j]``."""
        img = self.image()                       # GUI x = i, y = j (upwards)
        write_png_gray(filename, img.T[::-1, :])
        return img
