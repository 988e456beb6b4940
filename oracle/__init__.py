"""CPU oracle for the karman_taichi hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (NumPy, plus a small C/OpenMP port under
``oracle/c``) of the reference's arithmetic for the SIMPLE / BiCGSTAB / CG path.
It is the checker the CUDA path is compared against; it is never the product.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import or execute anything under ``oracle/``.
``karman_taichi_b200`` never imports it.

How the oracle is pinned: Taichi 0.9.0 cannot be installed in this
image (no wheel for Python 3.12, no network), so the reference itself cannot be run
here, and it is pure Python + Taichi (no C/C++ sources to compile into ``oracle/_ref``).
The reference keeps no golden vectors, no expected iteration counts and no field
snapshots.  PARITY WITH REFERENCE *OUTPUT* IS THEREFORE UNPINNED ("parity unpinned" at the
bit level: no number produced by the reference program itself is available to compare
with).  What the reference does hold -- one golden data file and a handful of
known-answer checks -- pins the oracle's behaviour and is re-expressed in ``tests/``:

* ``data/ghia1982.dat`` (Ghia et al. 1982 centreline profiles, compared at
  ``bicgsimp.py:291-304``)              -> ``tests/golden/ghia1982_re100.json``
* ``np.allclose`` against ``np.linalg.solve`` on a random 5-diagonal system
  (``debug/ti_bicg_test.py:45-98``)     -> ``tests/test_oracle_krylov.py``
* ``np.allclose(Ax, b)`` after CG (``benchmark/2dcg.py:129-132``)
* residual / convergence-history comparison against SciPy BiCGSTAB and a sparse
  direct solve (``benchmark/cg_to_sparse/test_cg_sparse.py:38-41``,
  ``validation.py:26-45``)

Two further anchors inside the repository: the NumPy restatement in serial-summation mode
(``oracle.krylov.set_summation('serial')``) and the C port on one thread are bit-identical
(``tests/test_oracle_cport.py``); and ``tests/golden/make_reference_golden.py`` is ready to execute the
reference's own source files under a serial stand-in for ``taichi`` and commit the outputs as
fixtures (``tests/test_reference_golden.py`` holds the oracle to them) -- NOT YET RUN: executing the
external sources needs the repository owner's approval, so the statement above stands.

Every function cites the reference file:line it follows.
"""
