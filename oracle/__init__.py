"""CPU oracle for the GlimSLib forward-simulation hot path.  TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy restatement of the discrete problem that the
reference (danielabler/glimslib) hands to FEniCS every time step:

* the weak form           ``glimslib/simulation/simulation_tumor_growth.py:110-124``
* the per-tissue variant  ``glimslib/simulation/simulation_tumor_growth_brain.py:24-125``
* the physics one-liners  ``glimslib/simulation_helpers/math_linear_elasticity.py:6-17,32-33``
                          ``glimslib/simulation_helpers/math_reaction_diffusion.py:2-3``
* the time loop           ``glimslib/simulation/simulation_base.py:236-317``
* (next scope row, N4)    the discrete adjoint of that loop for the controls and misfit of
                          ``image_based_optimization.py:660-700`` -- :mod:`oracle.adjoint`, checked against finite differences

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or as the
reported CPU baseline.  The product (``glimslib_b200``) never imports it.

PARITY UNPINNED (against FEniCS): the arithmetic of this path lives in
FEniCS 2017.2/2018.1 (UFL/FFC/DOLFIN/PETSc), which is neither vendored in the
reference tree nor installable here, and the reference's own tests hold no
golden vector for ``solver.solve()``.  What *is* pinned:

* the closed-form P1 element tensors in :mod:`oracle.fem` are checked against a
  literal quadrature evaluation of the reference's weak-form text
  (:mod:`oracle.weakform`) and against finite differences;
* closed-form solutions that P1 elements reproduce exactly -- free growth of a
  minimally supported body (stress-free dilatation ``u = c*gamma*(x - x_fixed)``)
  and the elasticity patch test -- are reproduced by the oracle to 1e-11
  (``tests/test_oracle.py``) and, independently of the oracle, by the CUDA path
  (``tests/test_gpu_analytic.py``);
* the structural known answers of the reference's unit tests
  (``test_unit_subDomains.py:36-74``, ``test_unit_boundaryConditions.py:90-108``)
  are reproduced in ``tests/test_dropin_host.py``, and the reference's unit tests
  themselves run unmodified in ``tests/test_reference_unit_tests_unmodified.py``.
"""
