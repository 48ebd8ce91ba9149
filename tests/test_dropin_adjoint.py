"""SURVEY.md 8f row N4 at the reference's API level: TumorGrowthBrain.misfit_gradient(controls, levels, targets, u_target) returns
the misfit of image_based_optimization.py:660-700 and its gradient with respect to the controls of run_for_adjoint
(simulation_tumor_growth_brain.py:127-145: D_WM, D_GM, rho_WM, rho_GM, coupling) -- what the reference gets from dolfin-adjoint's
ReducedFunctional -- through NonlinearVariationalSolver.adjoint_gradient -> glims_adjoint.

CPU part (host plumbing, oracle-backed stand-in engine): the gradient equals central finite differences of the misfit of the
drop-in's own forward runs (run_for_adjoint).  GPU part: the same call on the real engine gives the same J and gradient."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

LEVELS = [0.15, 0.45]
P_TRUE = np.array([0.10, 0.02, 0.15, 0.05, 0.10])
P = np.array([0.13, 0.03, 0.11, 0.07, 0.15])


def _sim():
    from glimslib_b200 import fenics_local as fenics
    from glimslib_b200.simulation.simulation_tumor_growth_brain import TumorGrowthBrain

    class Boundary(fenics.SubDomain):
        def inside(self, x, on_boundary):
            return on_boundary

    mesh = fenics.RectangleMesh(fenics.Point(0.0, 0.0), fenics.Point(4, 2), 10, 6)
    lab = fenics.MeshFunction("size_t", mesh, 2)
    lab.array()[:] = 1 + np.minimum((mesh.cell_midpoints()[:, 0] // 1).astype(int), 3)     # strips: CSF GM WM Ventricles
    sim = TumorGrowthBrain(mesh)
    sim.setup_global_parameters(subdomains=lab, domain_names={1: "CSF", 2: "GM", 3: "WM", 4: "Ventricles"},
                                boundaries={'all': Boundary()},
                                dirichlet_bcs={'clamp': {'bc_value': fenics.Constant((0.0, 0.0)), 'named_boundary': 'all', 'subspace_id': 0}})
    iv = {0: fenics.Constant((0.0, 0.0)), 1: fenics.Expression('0.8*exp(-3*(pow(x[0]-2.0,2)+pow(x[1]-1.0,2)))', degree=1)}
    sim.setup_model_parameters(iv_expression=iv, sim_time=3, sim_time_step=1, E_GM=3e-3, E_WM=3e-3, E_CSF=1e-3, E_VENT=1e-3,
                               nu_GM=0.4, nu_WM=0.4, nu_CSF=0.47, nu_VENT=0.3, D_GM=P_TRUE[1], D_WM=P_TRUE[0],
                               rho_GM=P_TRUE[3], rho_WM=P_TRUE[2], coupling=P_TRUE[4])
    return sim


def _final(sim, p, tmp_path):
    sim.run_for_adjoint(list(p), output_dir=str(tmp_path))
    return sim.solution.vector().get_local().reshape(-1, 3).copy()


def _targets(xN):
    th = lambda c, lv: 0.5 * (np.tanh((c - lv) / 0.01) + 1.0)
    return [th(xN[:, 2], lv) for lv in LEVELS], xN[:, :2].copy()


def _misfit(sim, x, tg, ut):
    """Mass-weighted misfit of a final state, evaluated with the drop-in's own assemble (independent of the engines)."""
    from glimslib_b200 import fenics_local as fenics
    V = fenics.FunctionSpace(sim.mesh, "Lagrange", 1)

    def sq(r):
        f = fenics.Function(V)
        f.vector().set_local(r)
        return fenics.assemble(f * f * fenics.dx)
    th = lambda c, lv: 0.5 * (np.tanh((c - lv) / 0.01) + 1.0)
    J = sum(sq(th(x[:, 2], lv) - t) for lv, t in zip(LEVELS, tg))
    return J + sum(sq(x[:, k] - ut[:, k]) for k in range(2))


def test_misfit_gradient_equals_finite_differences_of_forward_runs(monkeypatch, tmp_path):
    from oracle_engine import OracleEngine
    import glimslib_b200.backend.problem as problem
    monkeypatch.setattr(problem, "Engine", OracleEngine)
    sim = _sim()
    tg, ut = _targets(_final(sim, P_TRUE, tmp_path))
    J0, g0 = sim.misfit_gradient(list(P_TRUE), LEVELS, tg, ut)
    assert abs(J0) < 1e-16 and np.abs(g0).max() < 1e-7                         # the target parameters: a stationary zero
    J, g = sim.misfit_gradient(list(P), LEVELS, tg, ut)
    assert abs(J - _misfit(sim, sim.solution.vector().get_local().reshape(-1, 3), tg, ut)) <= 1e-10 * J
    fd = np.zeros(5)
    for k in range(5):
        h = 1e-5 * max(P[k], 1e-2)
        e = np.zeros(5)
        e[k] = h
        fd[k] = (_misfit(sim, _final(sim, P + e, tmp_path), tg, ut) - _misfit(sim, _final(sim, P - e, tmp_path), tg, ut)) / (2 * h)
    assert np.abs(g - fd).max() <= 2e-5 * np.abs(fd).max(), (g, fd)


@pytest.mark.gpu
def test_misfit_gradient_on_the_device_matches_the_oracle_engine(monkeypatch, tmp_path):
    from oracle_engine import OracleEngine
    import glimslib_b200.backend.problem as problem
    real = problem.Engine
    monkeypatch.setattr(problem, "Engine", OracleEngine)
    ref = _sim()
    tg, ut = _targets(_final(ref, P_TRUE, tmp_path))
    J_ref, g_ref = ref.misfit_gradient(list(P), LEVELS, tg, ut)
    monkeypatch.setattr(problem, "Engine", real)
    sim = _sim()
    J, g = sim.misfit_gradient(list(P), LEVELS, tg, ut)
    assert abs(J - J_ref) <= 1e-7 * abs(J_ref)
    assert np.abs(g - g_ref).max() <= 1e-5 * np.abs(g_ref).max(), (g, g_ref)


def _recover(sim, tmp_path):
    from scipy.optimize import minimize
    tg, ut = _targets(_final(sim, P_TRUE, tmp_path))
    history = []

    def fun(p):
        J, g = sim.misfit_gradient(list(p), LEVELS, tg, ut)
        history.append(J)
        return J, g
    p0 = np.array([0.14, 0.03, 0.11, 0.08, 0.16])
    res = minimize(fun, p0, jac=True, method="L-BFGS-B", bounds=[(1e-3, 0.5)] * 5, options=dict(maxiter=80, ftol=1e-18, gtol=1e-13))
    return history, res


def test_inverse_problem_recovers_the_parameters(monkeypatch, tmp_path):
    """The reference's production loop in miniature (image_based_optimization.py:703-767: scipy L-BFGS-B over the controls of
    run_for_adjoint with the dolfin-adjoint gradient): synthetic targets from known parameters, a perturbed start, box bounds;
    the misfit falls from 0.11 to round-off and all five controls come back."""
    from oracle_engine import OracleEngine
    import glimslib_b200.backend.problem as problem
    monkeypatch.setattr(problem, "Engine", OracleEngine)
    history, res = _recover(_sim(), tmp_path)
    assert history[0] > 0.05 and res.fun <= 1e-12 * history[0], (history[0], res.fun)
    assert np.abs(res.x - P_TRUE).max() <= 1e-5 * P_TRUE.max(), res.x


@pytest.mark.gpu
def test_inverse_problem_on_the_device(tmp_path):
    """The same loop with every forward run and every gradient on the GPU."""
    history, res = _recover(_sim(), tmp_path)
    assert history[0] > 0.05 and res.fun <= 1e-10 * history[0], (history[0], res.fun)
    assert np.abs(res.x - P_TRUE).max() <= 1e-4 * P_TRUE.max(), res.x
