"""Device fields: the stand-in for ``ti.field`` on the hot path.

The reference places dense Taichi fields of shape ``(NX, NY)`` (vectors) and ``(NX, NY, 5)``
(stencil coefficients, ``simple.py:42-47``) and reads them back with ``to_numpy()``.  Here a field is a
PyTorch CUDA tensor (PyTorch only owns the memory) in the layout the kernels want:

* vector  ``(NX, NY)``    -> one pitched plane ``(NX, pitch)``, ``pitch`` = NY rounded up to 512 bytes (``kt_pitch_for``),
  pad columns always zero;
* coef    ``(NX, NY, 5)`` -> five such planes ``(5, NX, pitch)`` (SoA: ap, aw, ae, an, as).

``to_numpy()`` / ``from_numpy()`` speak the reference's logical layout.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_TORCH = {np.dtype('float64'): torch.float64, np.dtype('float32'): torch.float32}
_CODE = {torch.float64: _lib.KT_F64, torch.float32: _lib.KT_F32}


def _device(device=None):
    if not torch.cuda.is_available():
        raise _lib.KtError('karman_taichi_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    if device is None:
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device(device)


class Field:
    """A ``(NX, NY)`` or ``(NX, NY, 5)`` real field resident in HBM."""

    def __init__(self, shape, dtype=np.float64, device=None):
        shape = tuple(int(s) for s in shape)
        if len(shape) not in (2, 3) or (len(shape) == 3 and shape[2] != 5):
            raise ValueError('Field shape must be (NX, NY) or (NX, NY, 5)')
        if shape[0] < 1 or shape[1] < 1:
            raise ValueError(f'Field shape {shape}: NX and NY must be at least 1 (an empty field has nothing to solve)')
        self.shape = shape
        self.np_dtype = np.dtype(dtype)
        self.dtype = _TORCH[self.np_dtype]
        self.dtype_code = _CODE[self.dtype]
        self.NX, self.NY = shape[0], shape[1]
        self.pitch = _lib.pitch_for(self.NY, self.dtype_code)
        self.device = _device(device)
        full = (5, self.NX, self.pitch) if len(shape) == 3 else (self.NX, self.pitch)
        self.data = torch.zeros(full, dtype=self.dtype, device=self.device)

    # -- reference-compatible surface -------------------------------------------------------
    @property
    def is_coef(self):
        return len(self.shape) == 3

    @property
    def ptr(self):
        return self.data.data_ptr()

    def to_numpy(self):
        if self.is_coef:
            return self.data[:, :, :self.NY].permute(1, 2, 0).contiguous().cpu().numpy()
        return self.data[:, :self.NY].contiguous().cpu().numpy()

    def from_numpy(self, arr):
        arr = np.ascontiguousarray(arr, dtype=self.np_dtype)
        if arr.shape != self.shape:
            raise ValueError(f'from_numpy: expected shape {self.shape}, got {arr.shape}')
        t = torch.from_numpy(arr).to(self.device)
        if self.is_coef:
            self.data[:, :, :self.NY] = t.permute(2, 0, 1)
        else:
            self.data[:, :self.NY] = t
        return self

    def fill(self, value):
        if self.is_coef:
            self.data[:, :, :self.NY] = value
        else:
            self.data[:, :self.NY] = value

    def copy_from(self, other):
        if other.shape != self.shape:
            raise ValueError('copy_from: shape mismatch')
        self.data.copy_(other.data)

    def _index(self, idx):
        """(i, j[, k]) checked against the LOGICAL shape: the pad columns [NY, pitch) must stay zero (every kernel sums and
        multiplies over whole padded rows), so they are not addressable."""
        if not isinstance(idx, tuple) or len(idx) != len(self.shape):
            raise IndexError(f'Field index must be {len(self.shape)} integers')
        out = []
        for v, n in zip(idx, self.shape):
            v = int(v)
            if not -n <= v < n:
                raise IndexError(f'Field index {idx} out of range for shape {self.shape}')
            out.append(v % n)
        return out

    def __getitem__(self, idx):
        if self.is_coef:
            i, j, k = self._index(idx)
            return float(self.data[k, i, j])
        i, j = self._index(idx)
        return float(self.data[i, j])

    def __setitem__(self, idx, value):
        if self.is_coef:
            i, j, k = self._index(idx)
            self.data[k, i, j] = value
        else:
            i, j = self._index(idx)
            self.data[i, j] = value


class ScalarField:
    """0-D field: ``alpha[None]`` (``cgsolver.py:22-24``).  Host-resident mirror of a solver scalar."""

    def __init__(self, value=0.0):
        self.value = float(value)
        self.shape = ()

    def __getitem__(self, idx):
        return self.value

    def __setitem__(self, idx, value):
        self.value = float(value)

    def to_numpy(self):
        return np.array(self.value)


def field_like(f, dtype=None):
    return Field(f.shape, f.np_dtype if dtype is None else dtype, f.device)
