"""Host side of ``fenics.project(expr, V)`` for the post-processing fields (``helper_classes.py:1566-1618``): the load vector
``int expr phi dx`` of an expression given at quadrature points, integrated with a collapsed Gauss-Legendre rule that is
exact for total polynomial degree <= 5; the consistent-mass solve itself runs on the device (``glims_mass_solve``).

The fields that are constant per cell or polynomial in the P1 concentration never come through here: the device integrates
those exactly in one kernel (``glims_project_fields``).  This path serves what the reference builds from already projected
functions (von Mises of the projected stress, the growth Jacobian of the projected expansion, the displacement norm)."""
import numpy as np


def simplex_quadrature(d, n=4):
    """(barycentric points [nq, d+1], weights [nq] summing to 1): exact for total degree <= 2n - 1 - (d - 1)."""
    g, w = np.polynomial.legendre.leggauss(n)
    g, w = 0.5 * (g + 1.0), 0.5 * w
    if d == 2:
        U, V = np.meshgrid(g, g, indexing="ij")
        WU, WV = np.meshgrid(w, w, indexing="ij")
        x, y = U.ravel(), (V * (1 - U)).ravel()
        wt = (WU * WV * (1 - U)).ravel() * 2.0
        lam = np.stack([1 - x - y, x, y], axis=1)
    else:
        U, V, W = np.meshgrid(g, g, g, indexing="ij")
        WU, WV, WW = np.meshgrid(w, w, w, indexing="ij")
        x, y, z = U.ravel(), (V * (1 - U)).ravel(), (W * (1 - U) * (1 - V)).ravel()
        wt = (WU * WV * WW * (1 - U) ** 2 * (1 - V)).ravel() * 6.0
        lam = np.stack([1 - x - y - z, x, y, z], axis=1)
    return lam, wt


def load_vector(mesh, integrand, n_comp=1, chunk=200000):
    """``load[v, k] = int integrand_k phi_v dx``.  ``integrand(cells_chunk [m, nb], lam [nq, nb], sl) -> [m, nq, n_comp]``
    (or [m, nq]) gives the expression at the quadrature points of the chunk of cells ``mesh.cells[sl]``."""
    d = mesh.dim
    lam, wt = simplex_quadrature(d)
    X = mesh.coords[mesh.cells]
    vol = np.abs(np.linalg.det(X[:, 1:] - X[:, :1])) / {2: 2.0, 3: 6.0}[d]
    load = np.zeros((mesh.num_vertices(), n_comp))
    for s in range(0, mesh.num_cells(), chunk):
        cc = mesh.cells[s:s + chunk]
        vals = np.asarray(integrand(cc, lam, slice(s, s + len(cc))), dtype=np.float64)
        if vals.ndim == 2:
            vals = vals[:, :, None]
        for a in range(d + 1):
            np.add.at(load, cc[:, a], vol[s:s + chunk, None] * np.einsum("q,eqk->ek", wt * lam[:, a], vals))
    return load
