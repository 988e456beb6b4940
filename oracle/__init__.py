"""CPU oracle for the karman_taichi hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (NumPy, plus a small C/OpenMP port under
``oracle/c``) of the reference's arithmetic for the SIMPLE / BiCGSTAB / CG path.
It is the checker the CUDA path is compared against; it is never the product.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import or execute anything under ``oracle/``.
``karman_taichi_b200`` never imports it.

How the oracle is pinned.  The reference is pure Python + Taichi 0.9.0 (no C/C++ sources to
compile into ``oracle/_ref``), and Taichi cannot be installed in this image (no wheel for Python 3.12, no network).  Its
arithmetic, however, lives entirely in its own ``.py`` files, so those files are EXECUTED UNMODIFIED under a serial
stand-in for the ``taichi`` module (``tests/golden/taichi_serial_shim.py``: fields = NumPy arrays, kernels = their Python
bodies run loop by loop in strict IEEE doubles) by ``tests/golden/make_reference_golden.py``, and what they produce is
committed as fixtures:

* ``tests/golden/reference_krylov.npz``   ``cgsolver.py`` / ``bicgsolver.py``: every intermediate vector, iterate, count
* ``tests/golden/reference_cavity.npz``   ``bicgsimp.py`` / ``simple.py``: coefficients, BCs, fields after outer steps
* ``tests/golden/reference_channel.npz``  ``karman_taichi.py``: matrices, Jacobi-BiCGSTAB iterates, fields after sub-iterations

``tests/test_reference_golden.py`` holds this oracle to them BIT FOR BIT in serial-summation mode (the order the stand-in's
``sum +=`` loops take); ``tests/test_gpu_reference_golden.py`` holds the CUDA path to the same fixtures.  Parity is therefore
pinned to the reference's own source, with one caveat stated plainly: the loops ran serially, not on Taichi's thread
pool -- the reference's own parallel ``sum +=`` is order-nondeterministic, so no run of it is "the" bit pattern either.

What the reference's own checks hold is re-expressed as well:

* ``data/ghia1982.dat`` (Ghia et al. 1982 centreline profiles, compared at ``bicgsimp.py:291-304``)
                                          -> ``tests/golden/ghia1982_re100.json``
* ``np.allclose`` against ``np.linalg.solve`` on a random 5-diagonal system (``debug/ti_bicg_test.py:45-98``)
* ``np.allclose(Ax, b)`` after CG (``benchmark/2dcg.py:129-132``)
* residual / convergence history against SciPy BiCGSTAB and a sparse direct solve
  (``benchmark/cg_to_sparse/test_cg_sparse.py:38-41``, ``validation.py:26-45``)

Summation orders (``oracle.krylov.set_summation``): 'pairwise' (NumPy), 'serial' (the stand-in's / a one-thread run's order;
bit-identical with the C port on one thread, ``tests/test_oracle_cport.py``) and the DEVICE order
(``oracle/device_order.py``: the exact order of the CUDA reductions from ``kt_reduce_plan`` / ``kt_resident_plan``), under
which the CUDA iterates and iteration counts are reproduced bit for bit (``tests/test_gpu_exact.py``).

Every function cites the reference file:line it follows.
"""
