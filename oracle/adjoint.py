"""Discrete adjoint of the backward-Euler time loop.  (oracle: test infrastructure for SURVEY.md section 8f, item N4)

The reference's production caller is an inverse problem: ``run_for_adjoint`` re-runs the forward model for new values of
(D_WM, D_GM, rho_WM, rho_GM, coupling) (``glimslib/simulation/simulation_tumor_growth_brain.py:127-145``) and
dolfin-adjoint differentiates a misfit of the final state -- thresholded concentration at two levels plus displacement --
with respect to those controls (``glimslib/optimization_workflow/image_based_optimization.py:660-700``; the smooth threshold
is ``0.5*(tanh((c - level)/0.01) + 1)``, ``:1404-1407``).

This module restates that gradient for the discrete problem of :mod:`oracle.fem` / :mod:`oracle.solver`:

    forward   R_n(x_n, x_{n-1}; p) = 0,   n = 1..N     (fem.assemble + Dirichlet rows)
    misfit    J(x_N) = sum_levels (th(c_N) - t_level)^T M (th(c_N) - t_level) + (u_N - u_t)^T (M x I_d) (u_N - u_t)
    adjoint   (dR_N/dx_N)^T l_N = -dJ/dx_N ;   (dR_n/dx_n)^T l_n = -(dR_{n+1}/dx_n)^T l_{n+1}
    gradient  dJ/dp = sum_n l_n^T dR_n/dp

M is the consistent P1 mass matrix (the reference L2-projects onto P1 before integrating; using the nodal values directly
is this restatement's choice and is stated here, it does not affect the adjoint structure).  dR_{n+1}/dx_n is minus the
mass block of the concentration rows.  The controls enter R linearly (D and rho per tissue, the coupling gamma through
beta = (2 mu + d lambda) gamma), so dR/dp is the difference of two residual assemblies.  The block-triangular Jacobian makes
the transposed solves a concentration solve after a displacement solve, i.e. the same two SPD solves as the forward step in
the opposite order -- which is what a device implementation would reuse.

PARITY UNPINNED: no dolfin-adjoint here and no golden gradient in the reference; the gradient is checked against central
finite differences of the forward run (tests/test_oracle_adjoint.py).
"""
from dataclasses import replace

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem, solver


def smooth_threshold(c, level, width=0.01):
    """image_based_optimization.py:1404-1407"""
    return 0.5 * (np.tanh((c - level) / width) + 1.0)


def mass_matrix(prob):
    """Consistent P1 mass matrix over the vertices."""
    d = prob.dim
    _, V = fem.geometry(prob.coords, prob.cells)
    Ml = fem.mass_local(d)
    n = d + 1
    rows = np.broadcast_to(prob.cells[:, :, None], (len(V), n, n)).ravel()
    cols = np.broadcast_to(prob.cells[:, None, :], (len(V), n, n)).ravel()
    vals = (V[:, None, None] * Ml[None]).ravel()
    M = sp.coo_matrix((vals, (rows, cols)), shape=(len(prob.coords),) * 2).tocsr()
    M.sum_duplicates()
    return M


def misfit(prob, x, targets, M=None):
    """targets: {'levels': {level: nodal target of the thresholded concentration}, 'u': nodal displacement target or None}.
    Returns (J, dJ/dx)."""
    d, nb = prob.dim, prob.dim + 1
    M = mass_matrix(prob) if M is None else M
    X = x.reshape(-1, nb)
    g = np.zeros_like(X)
    J = 0.0
    c = X[:, d]
    for level, tgt in targets.get("levels", {}).items():
        th = smooth_threshold(c, level)
        r = th - tgt
        Mr = M @ r
        J += r @ Mr
        dth = 0.5 * (1.0 - np.tanh((c - level) / 0.01) ** 2) / 0.01
        g[:, d] += 2.0 * Mr * dth
    if targets.get("u") is not None:
        ut = np.asarray(targets["u"]).reshape(-1, d)
        for k in range(d):
            r = X[:, k] - ut[:, k]
            Mr = M @ r
            J += r @ Mr
            g[:, k] += 2.0 * Mr
    return J, g.ravel()


def with_controls(prob, p, control_spec):
    """New Problem with the controls set.  control_spec: list of ('D'|'rho'|'gamma', material index or None for all)."""
    mats = prob.mats
    new = {k: getattr(mats, k).astype(np.float64).copy() for k in ("mu", "lam", "D", "rho", "gamma")}
    for val, (name, m) in zip(p, control_spec):
        if m is None:
            new[name][:] = val
        else:
            new[name][m] = val
    return replace(prob, mats=fem.Materials(**new))


def forward(prob, x0, n_steps, **newton_kw):
    kw = dict(linear="lu", rtol=1e-13, atol=1e-15)
    kw.update(newton_kw)
    geom = fem.geometry(prob.coords, prob.cells)
    xs = [x0.copy()]
    x = np.zeros_like(x0)
    for _ in range(n_steps):
        x, _ = solver.newton(prob, x, xs[-1], geom=geom, **kw)
        xs.append(x.copy())
    return xs


def residual(prob, x, x_prev, geom):
    F, _ = fem.assemble(prob, x, x_prev, False, geom)
    if len(prob.bc_dofs):
        F = F.copy()
        F[prob.bc_dofs] = x[prob.bc_dofs] - prob.bc_vals
    return F


def gradient(prob, x0, n_steps, targets, p, control_spec):
    """J and dJ/dp by one forward and one adjoint sweep."""
    pr = with_controls(prob, p, control_spec)
    geom = fem.geometry(pr.coords, pr.cells)
    xs = forward(pr, x0, n_steps)
    M = mass_matrix(pr)
    J, dJdx = misfit(pr, xs[-1], targets, M)
    d, nb = pr.dim, pr.dim + 1
    # dR_{n+1}/dx_n = -(mass block on the concentration dofs); Dirichlet rows are identities in x_{n+1} only
    ic = np.arange(len(pr.coords)) * nb + d
    B = sp.coo_matrix((-M.tocoo().data, (ic[M.tocoo().row], ic[M.tocoo().col])), shape=(pr.ndof, pr.ndof)).tocsr()
    if len(pr.bc_dofs):
        keep = np.ones(pr.ndof)
        keep[pr.bc_dofs] = 0.0
        B = sp.diags(keep) @ B
    grad = np.zeros(len(p))
    rhs = -dJdx
    for n in range(n_steps, 0, -1):
        F, Jn = fem.assemble(pr, xs[n], xs[n - 1], True, geom)
        _, Jn = fem.apply_dirichlet(pr, F, Jn, xs[n])
        lam = spla.spsolve(Jn.T.tocsc(), rhs)
        R0 = residual(pr, xs[n], xs[n - 1], geom)
        for k, (name, m) in enumerate(control_spec):
            q = np.array(p, dtype=np.float64)
            q[k] += 1.0                               # R is linear in every control: the unit difference is exact
            Rk = residual(with_controls(prob, q, control_spec), xs[n], xs[n - 1], geom)
            grad[k] += lam @ (Rk - R0)
        rhs = -(B.T @ lam)
    return J, grad


def functional(prob, x0, n_steps, targets, p, control_spec):
    pr = with_controls(prob, p, control_spec)
    xs = forward(pr, x0, n_steps)
    return misfit(pr, xs[-1], targets)[0]
