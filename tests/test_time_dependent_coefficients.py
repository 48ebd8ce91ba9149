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


def test_comparison_of_two_runs(monkeypatch, tmp_path):
    """helper_classes.Comparison (reference :1975-2035), as test_case_comparison_2D_atlas.py:203-210 uses it: TumorGrowth with
    per-tissue dict parameters against TumorGrowthBrain with the equivalent scalars -- zero difference -- and against a run with
    another diffusivity -- a difference carried by the concentration and, through the coupling, by the displacement."""
    from oracle_engine import OracleEngine
    import glimslib_b200.backend.problem as problem
    from glimslib_b200 import fenics_local as fenics
    from glimslib_b200.simulation.simulation_tumor_growth import TumorGrowth
    from glimslib_b200.simulation.simulation_tumor_growth_brain import TumorGrowthBrain
    from glimslib_b200.simulation_helpers.helper_classes import Comparison, AnyDimPoint
    monkeypatch.setattr(problem, "Engine", OracleEngine)

    class Boundary(fenics.SubDomain):
        def inside(self, x, on_boundary):
            return on_boundary

    mesh = fenics.RectangleMesh(AnyDimPoint((0.0, 0.0)), fenics.Point(4, 2), 8, 4)
    lab = fenics.MeshFunction("size_t", mesh, 2)
    lab.array()[:] = 1 + (mesh.cell_midpoints()[:, 0] // 1).astype(int)            # four vertical strips: CSF GM WM Ventricles
    tm = {1: "CSF", 2: "GM", 3: "WM", 4: "Ventricles"}
    bcs = {'clamp': {'bc_value': fenics.Constant((0.0, 0.0)), 'named_boundary': 'all', 'subspace_id': 0}}
    iv = {0: fenics.Constant((0.0, 0.0)), 1: fenics.Expression('exp(-4*(pow(x[0]-2.0,2)+pow(x[1]-1.0,2)))', degree=1)}

    def generic(d_wm):
        s = TumorGrowth(mesh)
        s.setup_global_parameters(subdomains=lab, domain_names=tm, boundaries={'all': Boundary()}, dirichlet_bcs=bcs)
        s.setup_model_parameters(iv_expression=iv, sim_time=3, sim_time_step=1,
                                 E={"CSF": 1e-3, "GM": 3e-3, "WM": 3e-3, "Ventricles": 1e-3},
                                 poisson={"CSF": 0.47, "GM": 0.4, "WM": 0.4, "Ventricles": 0.3},
                                 diffusion={"CSF": 0, "GM": 0.02, "WM": d_wm, "Ventricles": 0},
                                 proliferation={"CSF": 0, "GM": 0.05, "WM": 0.05, "Ventricles": 0},
                                 coupling={"CSF": 0.1, "GM": 0.1, "WM": 0.1, "Ventricles": 0.1})
        s.run(save_method=None, plot=False, output_dir=str(tmp_path))
        return s

    a = generic(0.1)
    b = TumorGrowthBrain(mesh)
    b.setup_global_parameters(subdomains=lab, domain_names=tm, boundaries={'all': Boundary()}, dirichlet_bcs=bcs)
    b.setup_model_parameters(iv_expression=iv, sim_time=3, sim_time_step=1, E_GM=3e-3, E_WM=3e-3, E_CSF=1e-3, E_VENT=1e-3,
                             nu_GM=0.4, nu_WM=0.4, nu_CSF=0.47, nu_VENT=0.3, D_GM=0.02, D_WM=0.1, rho_GM=0.05, rho_WM=0.05,
                             coupling=0.1)
    b.run(save_method=None, plot=False, output_dir=str(tmp_path))
    same = Comparison(a, b)
    assert same.shared_recording_steps == [0, 1, 2, 3]
    df = same.compare()
    assert list(df.columns) == ["recording_step", "errornorm", "errornorm_displacement", "errornorm_concentration"]
    assert df["errornorm"].max() <= 1e-14
    c = generic(0.2)
    diff = Comparison(a, c)
    df = diff.compare(slice(1, None))
    assert (df["errornorm_concentration"] > 1e-4).all() and (df["errornorm_displacement"] > 0).all()
    e = df.iloc[-1]
    assert abs(e["errornorm"] ** 2 - e["errornorm_displacement"] ** 2 - e["errornorm_concentration"] ** 2) <= 1e-12 * e["errornorm"] ** 2
    xa = a.results.get_result(3).get_field().vector().get_local()
    xc = c.results.get_result(3).get_field().vector().get_local()
    assert diff.compute_max_difference(3) == float(np.max(xa - xc))
    by = diff.get_difference_by_subspace(3)
    assert set(by) == {"displacement", "concentration"}
    assert np.allclose(by["concentration"].vector().get_local(), (xa - xc)[2::3])
