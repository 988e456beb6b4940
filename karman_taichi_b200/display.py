"""``Display`` -- post-processing / image dump of the solver fields (``display.py:4-49``; monolith
``post_process_field`` + ``ti.GUI`` frame, ``karman_taichi.py:544-583,662-676``).

Off the hot path: the fields are read back once per frame and processed with NumPy; the PNG is
written with zlib (no ti.GUI / matplotlib in this image).  Same arithmetic as the reference: cell-centre
interpolation ``0.5 (u[i,j] + u[i+1,j])``, ``0.5 (v[i,j] + v[i,j+1])``, per-field min/max scaling
``(f - min) / (max - min + 1e-9)``, the five panels (u, v, p, pcor, mdiv) concatenated along j; ``ti.GUI`` shows
``img[i, j]`` with i to the right and j upwards, which the PNG reproduces.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np


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


def scale_field(f):
    """``display.py:17-26``"""
    f_max, f_min = max(float(f.max()), -1.0e9), min(float(f.min()), 1.0e9)
    return (f - f_min) / (f_max - f_min + 1.0e-9)


class Display:
    def __init__(self, solver, *args):
        self.solver = solver
        self.nx, self.ny = solver.nx, solver.ny
        self.real = solver.real
        self.udisp = self.vdisp = self.pdisp = self.pcordisp = self.mdivdisp = None

    def post_process_field(self):
        """``display.py:28-40``"""
        s, nx, ny = self.solver, self.nx, self.ny
        u, v = s.u.to_numpy(), s.v.to_numpy()
        self.udisp = scale_field(0.5 * (u[0:nx + 2, :] + u[1:nx + 3, :]))
        self.vdisp = scale_field(0.5 * (v[:, 0:ny + 2] + v[:, 1:ny + 3]))
        self.pdisp = scale_field(s.p.to_numpy())
        self.pcordisp = scale_field(s.pcor.to_numpy())
        mdiv = getattr(s, 'mdiv', None) or getattr(s, 'udiv')
        self.mdivdisp = scale_field(mdiv.to_numpy())

    def image(self):
        self.post_process_field()
        return np.concatenate((self.udisp, self.vdisp, self.pdisp, self.pcordisp, self.mdivdisp), axis=1)

    def display(self, filename):
        """``display.py:42-49``: (nx+2) x 5(ny+2) frame saved to ``filename``."""
        img = self.image()                       # img[i, j]; GUI x = i, y = j (upwards)
        write_png_gray(filename, img.T[::-1, :])
        return img
