"""Closed-form P1 element tensors and scipy assembly.  (oracle: test infrastructure)

Restates, for P1 simplices with per-cell-constant coefficients, the residual
``F = F_m + F_rd`` and its Gateaux derivative ``J`` declared at
``glimslib/simulation/simulation_tumor_growth.py:110-124`` with the physics of
``glimslib/simulation_helpers/math_linear_elasticity.py:6-17,32-33`` and
``math_reaction_diffusion.py:2-3``.  All integrands are polynomials of degree
<= 3 on each cell, so FFC's quadrature is exact and these closed forms are the
same discrete equations up to round-off (checked against a literal quadrature
evaluation of the weak form in :mod:`oracle.weakform`).

Unknown layout used throughout the repo ("vertex-blocked"):
``dof(v, k) = v * (d + 1) + k`` with ``k < d`` the displacement components and
``k == d`` the concentration.
"""
from dataclasses import dataclass, field
from math import factorial

import numpy as np
import scipy.sparse as sp


def compute_mu(E, nu):
    """math_linear_elasticity.py:6-7"""
    return E / (2.0 * (1.0 + nu))


def compute_lambda(E, nu):
    """math_linear_elasticity.py:9-10"""
    return E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))


@dataclass
class Materials:
    """Per-label coefficient table keyed by *compact* material index
    (helper_classes.py:564-603 turns ``{tissue name: value}`` into a per-cell
    lookup; we key by label id)."""
    mu: np.ndarray
    lam: np.ndarray
    D: np.ndarray
    rho: np.ndarray
    gamma: np.ndarray

    @staticmethod
    def from_E_nu(E, nu, D, rho, gamma):
        E, nu, D, rho, gamma = (np.atleast_1d(np.asarray(a, dtype=np.float64)) for a in (E, nu, D, rho, gamma))
        return Materials(compute_mu(E, nu), compute_lambda(E, nu), D, rho, gamma)

    def table(self):
        return np.ascontiguousarray(np.stack([self.mu, self.lam, self.D, self.rho, self.gamma], axis=1))


@dataclass
class Problem:
    coords: np.ndarray          # (nv, d) f64
    cells: np.ndarray           # (nc, d+1) i32
    cell_mat: np.ndarray        # (nc,) i32 compact material index
    mats: Materials
    dt: float
    bc_dofs: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))
    bc_vals: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.float64))
    f_ext: np.ndarray = None    # (ndof,) external load: F = F_int(x) - f_ext

    @property
    def dim(self):
        return self.coords.shape[1]

    @property
    def ndof(self):
        return self.coords.shape[0] * (self.dim + 1)


def geometry(coords, cells):
    """Barycentric gradients G[e, a, :] = grad(lambda_a) and |K| (positive)."""
    d = coords.shape[1]
    X = coords[cells]                                   # (nc, d+1, d)
    Jm = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))   # columns = edge vectors
    det = np.linalg.det(Jm)
    Jinv = np.linalg.inv(Jm)                            # rows = grad lambda_1..d
    G = np.empty_like(X)
    G[:, 1:, :] = Jinv
    G[:, 0, :] = -Jinv.sum(axis=1)
    V = np.abs(det) / factorial(d)
    return G, V


def mass_local(d):
    """M_ab / |K| = (1 + delta_ab) / ((d+1)(d+2))"""
    n = d + 1
    return (np.ones((n, n)) + np.eye(n)) / ((d + 1) * (d + 2))


def triple_local(d):
    """T_abc / |K| = d! * prod(multiplicity!) / (d+3)!"""
    n = d + 1
    T = np.empty((n, n, n))
    for a in range(n):
        for b in range(n):
            for c in range(n):
                mult = np.bincount([a, b, c], minlength=n)
                T[a, b, c] = factorial(d) * np.prod([factorial(m) for m in mult]) / factorial(d + 3)
    return T


def element_tensors(prob, x, x_prev, want_jacobian=True, G=None, V=None):
    """Fe[e, a, k] and Ke[e, a, k, b, l] (k, l component indices, d = concentration)."""
    d = prob.dim
    n = d + 1
    if G is None:
        G, V = geometry(prob.coords, prob.cells)
    m = prob.cell_mat
    mu, lam, D, rho, gam = (getattr(prob.mats, k)[m] for k in ("mu", "lam", "D", "rho", "gamma"))
    dt = prob.dt
    xe = x.reshape(-1, n)[prob.cells]                   # (nc, n, n)
    ue, ce = xe[:, :, :d], xe[:, :, d]
    cpe = x_prev.reshape(-1, n)[prob.cells][:, :, d]
    Ml, Tl = mass_local(d), triple_local(d)
    beta = (2.0 * mu + d * lam) * gam                   # tr(sigma(v)) factor times coupling

    # --- residual -----------------------------------------------------------
    gradu = np.einsum("ebi,ebj->eij", ue, G)            # du_i/dx_j
    eps = 0.5 * (gradu + np.transpose(gradu, (0, 2, 1)))
    tr = np.trace(eps, axis1=1, axis2=2)
    sig = 2.0 * mu[:, None, None] * eps + (lam * tr)[:, None, None] * np.eye(d)[None]
    cbar = ce.mean(axis=1)
    Fe = np.empty((len(V), n, n))
    Fe[:, :, :d] = V[:, None, None] * (np.einsum("eij,eaj->eai", sig, G)
                                       - (beta * cbar)[:, None, None] * G)
    GG = np.einsum("eai,ebi->eab", G, G)
    Fc = V[:, None] * np.einsum("ab,eb->ea", Ml, ce - cpe)
    Fc += (dt * D * V)[:, None] * np.einsum("eab,eb->ea", GG, ce)
    Fc -= (dt * rho * V)[:, None] * (np.einsum("ab,eb->ea", Ml, ce)
                                     - np.einsum("abc,eb,ec->ea", Tl, ce, ce))
    Fe[:, :, d] = Fc
    if not want_jacobian:
        return Fe, None

    # --- Jacobian -----------------------------------------------------------
    Ke = np.zeros((len(V), n, n, n, n))
    I = np.eye(d)
    Kuu = (mu[:, None, None, None, None] * (GG[:, :, None, :, None] * I[None, None, :, None, :]
                                            + np.einsum("eaj,ebi->eaibj", G, G))
           + lam[:, None, None, None, None] * np.einsum("eai,ebj->eaibj", G, G))
    Ke[:, :, :d, :, :d] = V[:, None, None, None, None] * Kuu
    Ke[:, :, :d, :, d] = (-(beta * V) / n)[:, None, None, None] * G[:, :, :, None]
    Kcc = V[:, None, None] * Ml[None] + (dt * D * V)[:, None, None] * GG
    Kcc -= (dt * rho * V)[:, None, None] * (Ml[None] - 2.0 * np.einsum("abc,ec->eab", Tl, ce))
    Ke[:, :, d, :, d] = Kcc
    return Fe, Ke


def dof_index(cells, d):
    n = d + 1
    return (cells[:, :, None].astype(np.int64) * n + np.arange(n)[None, None, :])   # (nc, n, n)


def assemble(prob, x, x_prev, want_jacobian=True, geom=None):
    """Global residual (without Dirichlet rows applied) and CSR Jacobian."""
    d = prob.dim
    n = d + 1
    G, V = geom if geom is not None else geometry(prob.coords, prob.cells)
    Fe, Ke = element_tensors(prob, x, x_prev, want_jacobian, G, V)
    idx = dof_index(prob.cells, d)
    F = np.bincount(idx.ravel(), weights=Fe.ravel(), minlength=prob.ndof)
    if prob.f_ext is not None:
        F = F - prob.f_ext
    if not want_jacobian:
        return F, None
    rows = np.broadcast_to(idx[:, :, :, None, None], Ke.shape).ravel()
    cols = np.broadcast_to(idx[:, None, None, :, :], Ke.shape).ravel()
    J = sp.coo_matrix((Ke.ravel(), (rows, cols)), shape=(prob.ndof, prob.ndof)).tocsr()
    J.sum_duplicates()
    return F, J


def apply_dirichlet(prob, F, J, x):
    """DOLFIN ``DirichletBC::apply(A)`` / ``apply(b, x)`` as used by the Newton
    problem [MEM]: row zeroed, diagonal 1, ``F[i] = x[i] - g[i]`` (columns kept)."""
    if len(prob.bc_dofs) == 0:
        return F, J
    F = F.copy()
    F[prob.bc_dofs] = x[prob.bc_dofs] - prob.bc_vals
    if J is not None:
        keep = np.ones(prob.ndof)
        keep[prob.bc_dofs] = 0.0
        J = sp.diags(keep) @ J + sp.diags(1.0 - keep)
        J = J.tocsr()
    return F, J


def load_vector(prob, body_force=None, source=None, neumann=()):
    """External load of stg:112-113,119-120: body force ``b``, RD source ``s`` and
    Neumann terms ``sum_i g_i v ds(i)`` on exterior facets (helper_classes.py:861-908).

    ``neumann`` items: ``(facets[nf, d] vertex ids, owning cell[nf], subspace_id, value)``.
    Subspace 1 terms carry the factor ``dt * D(cell)`` (stg:120)."""
    d = prob.dim
    n = d + 1
    f = np.zeros((prob.coords.shape[0], n))
    _, V = geometry(prob.coords, prob.cells)
    if body_force is not None:
        b = np.asarray(body_force, dtype=np.float64)
        for a in range(n):
            np.add.at(f[:, :d], prob.cells[:, a], (V / n)[:, None] * b[None, :])
    if source is not None:
        for a in range(n):
            np.add.at(f[:, d], prob.cells[:, a], prob.dt * float(source) * V / n)
    for facets, owner, subspace, value in neumann:
        Xf = prob.coords[facets]
        if d == 2:
            area = np.linalg.norm(Xf[:, 1] - Xf[:, 0], axis=1)
        else:
            area = 0.5 * np.linalg.norm(np.cross(Xf[:, 1] - Xf[:, 0], Xf[:, 2] - Xf[:, 0]), axis=1)
        w = area / d
        if subspace == 0:
            g = np.asarray(value, dtype=np.float64)
            for k in range(d):
                np.add.at(f[:, :d], facets[:, k], w[:, None] * g[None, :])
        else:
            Dc = prob.mats.D[prob.cell_mat[owner]]
            for k in range(d):
                np.add.at(f[:, d], facets[:, k], prob.dt * Dc * float(value) * w)
    return f.ravel()
