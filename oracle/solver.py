"""Newton loop and backward-Euler time loop.  (oracle: test infrastructure)

Restates ``self.solver.solve()`` (``simulation_base.py:302``) as configured at
``simulation_tumor_growth.py:126-130`` -- DOLFIN ``NonlinearVariationalSolver``
with ``nonlinear_solver='snes'`` and every other option at its default [MEM]:
newtonls, basic (full-step) line search, rtol 1e-9, atol 1e-10, stol 1e-16,
max 50 iterations, error on non-convergence -- and the time loop of
``simulation_base.py:253-312`` (t=0 record, ``keep_nth``, stop at
``t <= T - 1e-5``, first Newton guess = fresh zero Function ``stg:100``).
"""
import numpy as np
import scipy.sparse.linalg as spla

from . import fem


class NotConverged(RuntimeError):
    pass


def newton(prob, x0, x_prev, rtol=1e-9, atol=1e-10, stol=1e-16, max_it=50,
           linear="lu", ksp_rtol=1e-5, geom=None, stats=None):
    """Solve F(x; x_prev) = 0.  ``linear``: 'lu' (tight) or 'gmres_ilu'
    (PETSc-default-like GMRES(30) + ILU(0), rtol 1e-5)."""
    x = x0.copy()
    if geom is None:
        geom = fem.geometry(prob.coords, prob.cells)
    f0 = None
    for it in range(max_it + 1):
        F, J = fem.assemble(prob, x, x_prev, True, geom)
        F, J = fem.apply_dirichlet(prob, F, J, x)
        fn = np.linalg.norm(F)
        if f0 is None:
            f0 = fn
        if stats is not None:
            stats.setdefault("fnorm", []).append(fn)
        if fn < atol or fn <= rtol * f0:
            return x, it
        if it == max_it:
            break
        if linear == "lu":
            dx = spla.spsolve(J.tocsc(), -F)
        else:
            ilu = spla.spilu(J.tocsc(), fill_factor=1.0, drop_tol=0.0)
            M = spla.LinearOperator(J.shape, ilu.solve)
            dx, info = spla.gmres(J, -F, M=M, restart=30, rtol=ksp_rtol, atol=0.0, maxiter=200)
            if info != 0:
                raise NotConverged("gmres info=%d" % info)
        x = x + dx
        if np.linalg.norm(dx) < stol * np.linalg.norm(x):
            return x, it + 1
    raise NotConverged("Newton: |F|=%g after %d iterations" % (fn, max_it))


def run(prob, x_init, sim_time, keep_nth=1, **newton_kw):
    """Returns (records, x_final); records = [(time, time_step, x copy)] starting
    with the t=0 record (simulation_base.py:271), as ``Results`` would hold them."""
    geom = fem.geometry(prob.coords, prob.cells)
    x_prev = x_init.copy()
    x = np.zeros_like(x_init)                 # stg:100 fresh Function
    records = [(0.0, 0, x_init.copy())]
    t, step = 0.0, 0
    while t <= sim_time - 1e-5:
        t += prob.dt
        step += 1
        try:
            x, _ = newton(prob, x, x_prev, geom=geom, **newton_kw)
        except NotConverged:
            break                             # simulation_base.py:303-305
        if step % keep_nth == 0:
            records.append((t, step, x.copy()))
        x_prev = x.copy()                     # simulation_base.py:312
    return records, x
