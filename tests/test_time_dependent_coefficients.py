"""Time-dependent coefficients through the drop-in: simulation_base._update_expressions(t) stamps `.t`
on parameters and von-Neumann values before every solve and the reference's UFL form re-evaluates them at assembly time
(simulation_tumor_growth.py:110-120).  The backend keeps the material table and the pre-integrated load vector on the
device, so `NonlinearVariationalSolver.solve()` has to re-push them when -- and only when -- they change.  Host logic only:
the device engine is replaced by the oracle-backed stand-in (tests/oracle_engine.py)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _sim(monkeypatch, flux_expr, tmp_path):
    from oracle_engine import OracleEngine
    import glimslib_b200.backend.problem as problem
    from glimslib_b200 import fenics_local as fenics
    from glimslib_b200.simulation.simulation_tumor_growth import TumorGrowth
    monkeypatch.setattr(problem, "Engine", OracleEngine)
    OracleEngine.instances.clear()

    class Boundary(fenics.SubDomain):
        def inside(self, x, on_boundary):
            return on_boundary

    mesh = fenics.RectangleMesh(fenics.Point(0, 0), fenics.Point(1, 1), 6, 6)
    sim = TumorGrowth(mesh)
    sim.setup_global_parameters(boundaries={'boundary_all': Boundary()},
                                dirichlet_bcs={'clamp': {'bc_value': fenics.Constant((0.0, 0.0)), 'named_boundary': 'boundary_all', 'subspace_id': 0}},
                                von_neumann_bcs={'flux': {'bc_value': flux_expr, 'named_boundary': 'boundary_all', 'subspace_id': 1}})
    iv = fenics.Expression('exp(-20*(pow(x[0]-0.5,2)+pow(x[1]-0.5,2)))', degree=1)
    sim.setup_model_parameters(iv_expression={0: fenics.Constant((0.0, 0.0)), 1: iv}, diffusion=0.01, coupling=0.1,
                               proliferation=0.1, E=1e-3, poisson=0.4, sim_time=3, sim_time_step=1)
    return sim, OracleEngine


def test_time_dependent_neumann_flux_is_re_integrated_every_step(monkeypatch, tmp_path):
    from glimslib_b200 import fenics_local as fenics
    flux = fenics.Expression('0.1*t', degree=1, t=0.0)
    sim, OE = _sim(monkeypatch, flux, tmp_path)
    sim.run(save_method=None, plot=False, output_dir=str(tmp_path))
    eng = OE.instances[-1]
    # the engine is configured at the first solve (t = 1), then gets a new load vector for t = 2 and t = 3
    assert eng.calls.count("set_load") == 3 and eng.calls.count("set_materials") == 1
    f = eng.f_ext.reshape(-1, 3)
    assert np.all(f[:, :2] == 0) and f[:, 2].sum() > 0
    # total flux integral at t = 3: dt * D * g * |boundary| = 1 * 0.01 * 0.3 * 4 (helper_classes.py:861-908; stg:120)
    assert abs(f[:, 2].sum() - 0.01 * 0.3 * 4.0) < 1e-12
    # a constant flux is integrated once
    sim2, OE = _sim(monkeypatch, fenics.Constant(0.2), tmp_path)
    sim2.run(save_method=None, plot=False, output_dir=str(tmp_path))
    assert OE.instances[-1].calls.count("set_load") == 1


def test_parameter_changed_between_runs_reaches_the_engine(monkeypatch, tmp_path):
    from glimslib_b200 import fenics_local as fenics
    sim, OE = _sim(monkeypatch, fenics.Constant(0.0), tmp_path)
    sim.run(save_method=None, plot=False, output_dir=str(tmp_path))
    c1 = sim.solution.vector().get_local()[2::3].sum()
    eng = OE.instances[-1]
    n = eng.calls.count("set_materials")
    sim.params.proliferation = 0.0            # inverse-problem call pattern: same mesh, new parameters
    sim.run(save_method=None, plot=False, output_dir=str(tmp_path))
    assert OE.instances[-1] is eng and eng.calls.count("set_materials") == n + 1
    assert eng.table[0, 3] == 0.0
    assert sim.solution.vector().get_local()[2::3].sum() < c1
