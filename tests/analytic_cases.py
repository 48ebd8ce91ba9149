"""Closed-form solutions of the coupled model that P1 elements reproduce to round-off on any mesh -- known answers that do not
come from the oracle (tests/test_oracle.py checks the oracle against them, tests/test_gpu_analytic.py the CUDA path).

* free growth: uniform concentration c, no proliferation, no diffusion, a body held only against rigid motion.  The growth strain
  c*gamma*I (math_linear_elasticity.compute_growth_induced_strain; the -sigma(v):(c gamma I) term of stg:110-114) is compatible,
  so the exact solution is the dilatation u = c*gamma*(x - x_fixed) with zero stress, whatever E and nu are (two stiffnesses are
  used).  Pins sign and magnitude of the coupling term.
* patch test: homogeneous material, no growth, boundary displacements of a linear field u = A x + b: the interior solution is
  that field (constant strain satisfies equilibrium)."""
import numpy as np

from oracle import fem, meshes


def _mesh(d, rng, jitter):
    if d == 2:
        coords, cells = meshes.rectangle_mesh((0, 0), (2.0, 1.0), 5, 4)
    else:
        coords, cells = meshes.box_mesh((0, 0, 0), (1.0, 1.4, 0.8), 3, 3, 2)
    bv = meshes.boundary_vertices(cells, len(coords))
    inner = np.setdiff1d(np.arange(len(coords)), bv)
    coords = coords.copy()
    coords[inner] += jitter * (rng.random((len(inner), d)) - 0.5)
    return coords, cells, bv


def free_growth_case(d, gamma=0.15, c=0.6):
    """Returns (problem, x0, exact displacement [n_v, d], c)."""
    rng = np.random.default_rng(5 + d)
    coords, cells, _ = _mesh(d, rng, 0.04)
    nb = d + 1
    mats = fem.Materials.from_E_nu([3e-3, 1e-3], [0.45, 0.3], [0.0, 0.0], [0.0, 0.0], [gamma, gamma])
    cell_mat = (coords[cells].mean(axis=1)[:, 0] > coords[:, 0].mean()).astype(np.int32)
    # statically determinate supports: one vertex fixed, and just enough components of other vertices to stop rotation,
    # each prescribed with the value the dilatation takes there
    v0 = int(np.argmin(np.linalg.norm(coords, axis=1)))
    far_x = int(np.argmax(coords[:, 0] - 10 * np.abs(coords[:, 1:]).sum(axis=1)))           # on the x axis
    exact = c * gamma * (coords - coords[v0])
    sup = [(v0, k) for k in range(d)] + [(far_x, 1)]
    if d == 3:
        far_y = int(np.argmax(coords[:, 1] - 10 * (np.abs(coords[:, 0]) + np.abs(coords[:, 2]))))
        sup += [(far_x, 2), (far_y, 2)]
    dofs = np.array([v * nb + k for v, k in sup], dtype=np.int64)
    vals = np.array([exact[v, k] for v, k in sup])
    order = np.argsort(dofs)
    prob = fem.Problem(coords, cells, cell_mat, mats, dt=1.0, bc_dofs=dofs[order], bc_vals=vals[order])
    x0 = np.zeros(len(coords) * nb)
    x0[d::nb] = c
    return prob, x0, exact, c


def patch_case(d):
    """Returns (problem, exact displacement [n_v, d])."""
    rng = np.random.default_rng(11 + d)
    coords, cells, bv = _mesh(d, rng, 0.05)
    nb = d + 1
    A, b = 0.01 * rng.standard_normal((d, d)), 0.01 * rng.standard_normal(d)
    exact = coords @ A.T + b
    dofs = np.sort((bv[:, None] * nb + np.arange(d)[None, :]).ravel()).astype(np.int64)
    vals = exact[dofs // nb, dofs % nb]
    mats = fem.Materials.from_E_nu([2e-3], [0.35], [0.0], [0.0], [0.0])
    prob = fem.Problem(coords, cells, np.zeros(len(cells), np.int32), mats, dt=1.0, bc_dofs=dofs, bc_vals=vals)
    return prob, exact
