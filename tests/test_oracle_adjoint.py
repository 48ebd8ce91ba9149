"""Oracle for the next scope row: the discrete adjoint of the time loop reproduces central finite
differences of the forward run.  No GPU, no dolfin-adjoint; this pins the restatement's own consistency only."""
import numpy as np
import pytest

from oracle import adjoint, fem, meshes


def _problem(d):
    if d == 2:
        coords, cells = meshes.rectangle_mesh((-1, -1), (1, 1), 8, 8)
    else:
        coords, cells = meshes.box_mesh((0, 0, 0), (1, 1, 1), 4, 4, 4)
    cen = coords[cells].mean(axis=1)
    cm = (cen[:, 0] > coords[:, 0].mean()).astype(np.int32)             # two tissues ("WM", "GM")
    mats = fem.Materials.from_E_nu([3e-3, 3e-3], [0.45, 0.40], [0.10, 0.02], [0.15, 0.05], [0.10, 0.10])
    bv = meshes.boundary_vertices(cells, len(coords))
    nb = d + 1
    dofs = np.sort((bv[:, None] * nb + np.arange(d)[None, :]).ravel()).astype(np.int64)
    prob = fem.Problem(coords, cells, cm, mats, dt=1.0, bc_dofs=dofs, bc_vals=np.zeros(len(dofs)))
    x0 = np.zeros(prob.ndof)
    ctr = coords.mean(axis=0) + 0.1
    x0[d::nb] = 0.8 * np.exp(-6.0 * ((coords - ctr) ** 2).sum(axis=1))
    return prob, x0


@pytest.mark.parametrize("d", [2, 3])
def test_adjoint_gradient_matches_finite_differences(d):
    """Controls as in run_for_adjoint (simulation_tumor_growth_brain.py:127-145): D and rho of two tissues, one coupling."""
    prob, x0 = _problem(d)
    spec = [("D", 0), ("D", 1), ("rho", 0), ("rho", 1), ("gamma", None)]
    p_true = np.array([0.10, 0.02, 0.15, 0.05, 0.10])
    n_steps = 3
    # synthetic targets from the "true" parameters, then evaluate the gradient away from them
    xs = adjoint.forward(adjoint.with_controls(prob, p_true, spec), x0, n_steps)
    nb = d + 1
    cN = xs[-1][d::nb]
    targets = {"levels": {0.12: adjoint.smooth_threshold(cN, 0.12), 0.4: adjoint.smooth_threshold(cN, 0.4)},
               "u": xs[-1].reshape(-1, nb)[:, :d].copy()}
    p = np.array([0.14, 0.03, 0.11, 0.08, 0.16])
    J, g = adjoint.gradient(prob, x0, n_steps, targets, p, spec)
    assert J > 0
    for k in range(len(p)):
        h = 1e-5 * max(abs(p[k]), 1e-2)
        pp, pm = p.copy(), p.copy()
        pp[k] += h
        pm[k] -= h
        fd = (adjoint.functional(prob, x0, n_steps, targets, pp, spec) - adjoint.functional(prob, x0, n_steps, targets, pm, spec)) / (2 * h)
        assert abs(g[k] - fd) <= 1e-5 * max(abs(fd), np.abs(g).max() * 1e-3), (k, g[k], fd)


def test_misfit_vanishes_at_the_target_parameters():
    prob, x0 = _problem(2)
    spec = [("D", 0), ("rho", 0)]
    p = np.array([0.10, 0.15])
    xs = adjoint.forward(adjoint.with_controls(prob, p, spec), x0, 2)
    targets = {"levels": {0.2: adjoint.smooth_threshold(xs[-1][2::3], 0.2)}, "u": xs[-1].reshape(-1, 3)[:, :2].copy()}
    J, g = adjoint.gradient(prob, x0, 2, targets, p, spec)
    assert J < 1e-20 and np.abs(g).max() < 1e-9
