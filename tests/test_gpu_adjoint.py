"""SURVEY.md 8f row N4 on the device: glims_adjoint (forward steps with the trajectory kept in HBM, misfit of the final state,
one backward sweep of transposed block solves, per-material gradient reductions) against oracle/adjoint.py -- the discrete
adjoint restated in numpy/scipy, itself checked against central finite differences of the forward run
(tests/test_oracle_adjoint.py).  Controls as in run_for_adjoint (simulation_tumor_growth_brain.py:127-145): D and rho of two
tissues and one coupling; misfit as in image_based_optimization.py:660-700 (two threshold levels + displacement)."""
import numpy as np
import pytest

from oracle import adjoint, fem, meshes

pytestmark = pytest.mark.gpu


def _problem(d):
    if d == 2:
        coords, cells = meshes.rectangle_mesh((-1, -1), (1, 1), 8, 8)
    else:
        coords, cells = meshes.box_mesh((0, 0, 0), (1, 1, 1), 4, 4, 4)
    cen = coords[cells].mean(axis=1)
    cm = (cen[:, 0] > coords[:, 0].mean()).astype(np.int32)
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
def test_device_adjoint_gradient_matches_the_oracle(d):
    from glimslib_b200.engine import Engine
    prob, x0 = _problem(d)
    nb = d + 1
    spec = [("D", 0), ("D", 1), ("rho", 0), ("rho", 1), ("gamma", None)]
    p_true = np.array([0.10, 0.02, 0.15, 0.05, 0.10])
    n_steps = 3
    xs = adjoint.forward(adjoint.with_controls(prob, p_true, spec), x0, n_steps)
    cN = xs[-1][d::nb]
    levels = [0.12, 0.4]
    targets = {"levels": {lv: adjoint.smooth_threshold(cN, lv) for lv in levels}, "u": xs[-1].reshape(-1, nb)[:, :d].copy()}
    p = np.array([0.14, 0.03, 0.11, 0.08, 0.16])
    J_ref, g_ref = adjoint.gradient(prob, x0, n_steps, targets, p, spec)
    pr = adjoint.with_controls(prob, p, spec)
    eng = Engine(pr.coords, pr.cells, pr.cell_mat)
    eng.set_materials(pr.mats.table())
    eng.set_dt(pr.dt)
    eng.set_dirichlet(pr.bc_dofs, pr.bc_vals)
    eng.set_prev(x0)
    eng.set_state(np.zeros(pr.ndof))
    J, grad = eng.adjoint_gradient(n_steps, levels, np.stack([targets["levels"][lv] for lv in levels]), targets["u"],
                                   snes_rtol=1e-13, snes_atol=1e-16, ksp_rtol=1e-13)
    # the forward run inside the call ends where the oracle's does
    x = eng.get_state()
    assert np.linalg.norm(x - adjoint.forward(pr, x0, n_steps)[-1]) <= 1e-9 * np.linalg.norm(x)
    assert abs(J - J_ref) <= 1e-8 * abs(J_ref)
    g = np.array([grad[0, 0], grad[1, 0], grad[0, 1], grad[1, 1], grad[:, 2].sum()])
    assert np.abs(g - g_ref).max() <= 1e-6 * np.abs(g_ref).max(), (g, g_ref)
    eng.close()


def test_device_adjoint_vanishes_at_the_target():
    from glimslib_b200.engine import Engine
    prob, x0 = _problem(2)
    xs = adjoint.forward(prob, x0, 2)
    c2 = xs[-1][2::3]
    eng = Engine(prob.coords, prob.cells, prob.cell_mat)
    eng.set_materials(prob.mats.table())
    eng.set_dt(prob.dt)
    eng.set_dirichlet(prob.bc_dofs, prob.bc_vals)
    eng.set_prev(x0)
    eng.set_state(np.zeros(prob.ndof))
    J, grad = eng.adjoint_gradient(2, [0.2], adjoint.smooth_threshold(c2, 0.2)[None, :], xs[-1].reshape(-1, 3)[:, :2],
                                   snes_rtol=1e-13, snes_atol=1e-16, ksp_rtol=1e-13)
    assert J < 1e-16 and np.abs(grad).max() < 1e-7
    eng.close()
