"""karman_taichi_b200 -- the SIMPLE / BiCGSTAB / CG hot path of houkensjtu/karman_taichi on B200.

Host classes keep the reference's names and signatures (``CGSolver``, ``BICGSolver``, ``SIMPLESolver``);
all arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of ``include/karman_b200.h``
(``libkarman_b200.so``, built in-tree by ``karman_taichi_b200.build``).  PyTorch only owns device memory and
streams.  There is no CPU fallback.
"""
from . import _lib
from ._lib import KtError
from .field import Field, ScalarField
from .cgsolver import CGSolver
from .bicgsolver import BICGSolver
from .simple import SIMPLESolver, lid_driven_cavity
from .karman import KarmanSolver
from .display import Display
from .jacobi import JacobiSolver
from .slab_karman import SlabKarmanSolver

__all__ = ['Field', 'ScalarField', 'CGSolver', 'BICGSolver', 'SIMPLESolver', 'lid_driven_cavity', 'KarmanSolver', 'Display', 'JacobiSolver', 'SlabKarmanSolver', 'KtError']
__version__ = '0.1.0'
