"""L2 projections of the derived fields onto P1, literally as the reference does them.  (oracle: test infrastructure)

``PostProcessTumorGrowth`` (``helper_classes.py:1566-1618, 1736-1786``) calls ``fenics.project(expr, V)`` for every field:
solve ``M q = int expr phi dx`` with the consistent P1 mass matrix.  Here every integral is evaluated by a collapsed
(Duffy) Gauss-Legendre rule that is exact for polynomials of total degree <= 5 -- enough for every polynomial integrand on
this path (degree <= 4) -- directly from the UFL text of ``math_linear_elasticity.py:12-71`` and
``math_reaction_diffusion.py:2-3``, and ``M`` is assembled by the same quadrature; the linear solves are sparse LU.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def simplex_quadrature(d, n=4):
    """(barycentric points [nq, d+1], weights [nq] summing to 1) of the collapsed Gauss-Legendre rule with n points per
    direction on the reference simplex: exact for total degree <= 2n - 1 - (d - 1)."""
    g, w = np.polynomial.legendre.leggauss(n)
    g, w = 0.5 * (g + 1.0), 0.5 * w                     # [0, 1]
    if d == 2:
        U, V = np.meshgrid(g, g, indexing="ij")
        WU, WV = np.meshgrid(w, w, indexing="ij")
        x, y = U.ravel(), (V * (1 - U)).ravel()
        wt = (WU * WV * (1 - U)).ravel() * 2.0            # reference triangle has area 1/2
        lam = np.stack([1 - x - y, x, y], axis=1)
    else:
        U, V, W = np.meshgrid(g, g, g, indexing="ij")
        WU, WV, WW = np.meshgrid(w, w, w, indexing="ij")
        x, y, z = U.ravel(), (V * (1 - U)).ravel(), (W * (1 - U) * (1 - V)).ravel()
        wt = (WU * WV * WW * (1 - U) ** 2 * (1 - V)).ravel() * 6.0
        lam = np.stack([1 - x - y - z, x, y, z], axis=1)
    return lam, wt


def _geometry(coords, cells):
    X = coords[cells]
    d = coords.shape[1]
    J = np.transpose(X[:, 1:] - X[:, :1], (0, 2, 1))              # columns = edge vectors
    vol = np.abs(np.linalg.det(J)) / {2: 2.0, 3: 6.0}[d]
    Jinv = np.linalg.inv(J)
    g = np.zeros((len(cells), d + 1, d))
    g[:, 1:, :] = Jinv                                             # grad lambda_a, a >= 1 = rows of J^-1
    g[:, 0, :] = -Jinv.sum(axis=1)
    return vol, g


def mass_matrix(coords, cells):
    d = coords.shape[1]
    nb = d + 1
    vol, _ = _geometry(coords, cells)
    lam, wt = simplex_quadrature(d)
    local = np.einsum("q,qa,qb->ab", wt, lam, lam)                 # int lambda_a lambda_b / |K|
    rows = np.repeat(cells, nb, axis=1).ravel()
    cols = np.tile(cells, (1, nb)).ravel()
    vals = (vol[:, None, None] * local[None]).ravel()
    n = len(coords)
    return sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsc()


def project(coords, cells, integrand):
    """``fenics.project``: integrand(lam [nq, nb]) -> values [n_cells, nq, k] at the quadrature points; returns [n_v, k]."""
    d = coords.shape[1]
    vol, _ = _geometry(coords, cells)
    lam, wt = simplex_quadrature(d)
    vals = integrand(lam)
    if vals.ndim == 2:
        vals = vals[:, :, None]
    k = vals.shape[2]
    load = np.zeros((len(coords), k))
    for a in range(d + 1):
        np.add.at(load, cells[:, a], vol[:, None] * np.einsum("q,eqk->ek", wt * lam[:, a], vals))
    lu = spla.splu(mass_matrix(coords, cells))
    return np.stack([lu.solve(load[:, j]) for j in range(k)], axis=1)


def derived_fields(coords, cells, cell_mat, table, x):
    """Every field of PostProcessTumorGrowth for the state x (vertex-blocked), projected as the reference projects them.
    table[m] = (mu, lambda, D, rho, gamma).  Returns a dict of [n_v, ...] arrays."""
    d = coords.shape[1]
    nb = d + 1
    nv, nc = len(coords), len(cells)
    X = x.reshape(nv, nb)
    u, c = X[:, :d], X[:, d]
    vol, g = _geometry(coords, cells)
    mu, lmb, rho, gam = (table[cell_mat, k] for k in (0, 1, 3, 4))
    gu = np.einsum("eai,eaj->eij", u[cells], g)                    # grad u [e, i, j] = d u_i / d x_j (constant per cell)
    eps = 0.5 * (gu + np.transpose(gu, (0, 2, 1)))                 # mle:12-13
    I = np.eye(d)
    sig = 2.0 * mu[:, None, None] * eps + lmb[:, None, None] * np.trace(eps, axis1=1, axis2=2)[:, None, None] * I   # mle:15-17
    const = lambda a: (lambda lam: np.repeat(a.reshape(nc, 1, -1), len(lam), axis=1))
    out = {}
    out["strain"] = project(coords, cells, const(eps)).reshape(nv, d, d)
    out["stress"] = project(coords, cells, const(sig)).reshape(nv, d, d)
    sh = out["stress"]
    # pressure / von Mises are built from the PROJECTED stress function and projected again (helper_classes.py:1586-1602)
    out["pressure"] = project(coords, cells, lambda lam: np.einsum("qa,eaii->eq", lam, sh[cells]) / 3.0)[:, 0]    # mle:19-21

    def vm(lam):
        s = np.einsum("qa,eaij->eqij", lam, sh[cells])
        dev = s - (np.trace(s, axis1=2, axis2=3) / 3.0)[:, :, None, None] * I                                     # mle:35-36
        return np.sqrt(1.5 * np.einsum("eqij,eqij->eq", dev, dev))                                                # mle:38-40
    out["von_mises"] = project(coords, cells, vm)[:, 0]
    out["total_jacobian"] = project(coords, cells, const(np.linalg.det(I + gu)))[:, 0]                            # mle:26-27
    cq = lambda lam: np.einsum("qa,ea->eq", lam, c[cells])
    mech = project(coords, cells, lambda lam: cq(lam) * gam[:, None])[:, 0]      # scalar factor of c * gamma * I (mle:32-33)
    out["mech_expansion"] = mech
    out["growth_jacobian"] = project(coords, cells, lambda lam: (1.0 + np.einsum("qa,ea->eq", lam, mech[cells])) ** d)[:, 0]   # mle:29-30
    out["logistic_growth"] = project(coords, cells, lambda lam: rho[:, None] * cq(lam) * (1.0 - cq(lam)))[:, 0]   # mrd:2-3
    # concentration in the deformed configuration, from the raw solution (helper_classes.py:1779-1786; mle:67-71):
    # c * det(I + c gamma I) / det(I + grad u)
    jt = np.linalg.det(I + gu)
    out["concentration_deformed"] = project(coords, cells, lambda lam: cq(lam) * (1.0 + gam[:, None] * cq(lam)) ** d / jt[:, None])[:, 0]
    out["displacement_norm"] = project(coords, cells, lambda lam: np.sqrt((np.einsum("qa,eai->eqi", lam, u[cells]) ** 2).sum(axis=2)))[:, 0]
    return out
