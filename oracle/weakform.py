"""Literal quadrature evaluation of the reference's weak form.  (oracle: test infrastructure)

Independent of :mod:`oracle.fem`'s closed forms: evaluates, cell by cell and
quadrature point by quadrature point, exactly the integrands written at
``glimslib/simulation/simulation_tumor_growth.py:110-120``

    F_m  = inner(sigma(u), eps(v0)) dx - inner(sigma(v0), c*gamma*I) dx
    F_rd = c v1 dx + dt D inner(grad c, grad v1) dx - c_prev v1 dx
           - dt rho c (1 - c/1.0) v1 dx

with ``sigma(w) = 2 mu sym(grad w) + lambda tr(sym(grad w)) I_d``
(``math_linear_elasticity.py:12-17``), growth strain ``c*gamma*I_d`` (``:32-33``)
and logistic growth (``math_reaction_diffusion.py:2-3``).  The Jacobian is the
complex-step-free central difference of this residual.  Pure-Python loops:
small meshes only.
"""
import numpy as np
from scipy.special import roots_jacobi


def simplex_rule(d, npts=4):
    """Collapsed Gauss-Jacobi rule on the unit simplex, exact to degree 2*npts-1."""
    if d == 2:
        xa, wa = roots_jacobi(npts, 0, 0)
        xb, wb = roots_jacobi(npts, 1, 0)
        pts, wts = [], []
        for i in range(npts):
            for j in range(npts):
                r, s = (xa[i] + 1) / 2, (xb[j] + 1) / 2
                pts.append((r * (1 - s), s))
                wts.append(wa[i] * wb[j] / 8.0)
        return np.array(pts), np.array(wts)
    xa, wa = roots_jacobi(npts, 0, 0)
    xb, wb = roots_jacobi(npts, 1, 0)
    xc, wc = roots_jacobi(npts, 2, 0)
    pts, wts = [], []
    for i in range(npts):
        for j in range(npts):
            for k in range(npts):
                r, s, t = (xa[i] + 1) / 2, (xb[j] + 1) / 2, (xc[k] + 1) / 2
                pts.append((r * (1 - s) * (1 - t), s * (1 - t), t))
                wts.append(wa[i] * wb[j] * wc[k] / 64.0)
    return np.array(pts), np.array(wts)


def residual(coords, cells, cell_mat, table, dt, x, x_prev):
    """Global residual by literal quadrature. ``table[m] = (mu, lam, D, rho, gamma)``."""
    d = coords.shape[1]
    n = d + 1
    pts, wts = simplex_rule(d)
    F = np.zeros(coords.shape[0] * n)
    xb = x.reshape(-1, n)
    xpb = x_prev.reshape(-1, n)
    I = np.eye(d)
    for e, cell in enumerate(cells):
        X = coords[cell]
        Jm = (X[1:] - X[0]).T
        detJ = abs(np.linalg.det(Jm))
        Jinv = np.linalg.inv(Jm)
        gref = np.vstack([-np.ones(d), np.eye(d)])       # grad of reference basis
        g = gref @ Jinv                                  # physical gradients (n, d)
        mu, lam, D, rho, gam = table[cell_mat[e]]
        u = xb[cell, :d]
        c = xb[cell, d]
        cp = xpb[cell, d]

        def sigma(gradw):
            eps = 0.5 * (gradw + gradw.T)
            return 2.0 * mu * eps + lam * np.trace(eps) * I

        grad_u = u.T @ g
        grad_c = c @ g
        for p, w in zip(pts, wts):
            phi = np.concatenate([[1.0 - p.sum()], p])
            cq, cpq = phi @ c, phi @ cp
            wq = w * detJ
            for a in range(n):
                for i in range(d):
                    grad_v = np.outer(I[i], g[a])        # grad of v0 = e_i phi_a
                    eps_v = 0.5 * (grad_v + grad_v.T)
                    val = np.sum(sigma(grad_u) * eps_v) - np.sum(sigma(grad_v) * (cq * gam * I))
                    F[cell[a] * n + i] += wq * val
                v1, grad_v1 = phi[a], g[a]
                val = (cq * v1 + dt * D * (grad_c @ grad_v1) - cpq * v1
                       - dt * (rho * cq * (1.0 - cq / 1.0)) * v1)
                F[cell[a] * n + d] += wq * val
    return F


def jacobian_fd(coords, cells, cell_mat, table, dt, x, x_prev, h=1e-6):
    """Dense central-difference Jacobian of :func:`residual` (tiny meshes only).
    The residual is at most quadratic in x, so central differences are exact up
    to round-off."""
    N = len(x)
    J = np.zeros((N, N))
    for j in range(N):
        e = np.zeros(N)
        e[j] = h
        J[:, j] = (residual(coords, cells, cell_mat, table, dt, x + e, x_prev)
                   - residual(coords, cells, cell_mat, table, dt, x - e, x_prev)) / (2 * h)
    return J
