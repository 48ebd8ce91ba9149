"""GPU test of the drop-in API: the reference's 2D two-subdomain script
(test_cases/test_simulation_tumor_growth/test_case_simulation_tumor_growth_2D_subdomains.py) re-typed against
glimslib_b200, every recorded step compared with the oracle on the same labels / IC / BCs."""
import os

import numpy as np
import pytest

from oracle import fem, solver as osolver

pytestmark = pytest.mark.gpu


def _script(tmp_path, save_method):
    from glimslib_b200 import fenics_local as fenics
    from glimslib_b200.simulation.simulation_tumor_growth import TumorGrowth

    class Boundary(fenics.SubDomain):
        def inside(self, x, on_boundary):
            return on_boundary

    nx = ny = 50
    mesh = fenics.RectangleMesh(fenics.Point(-5, -5), fenics.Point(5, 5), nx, ny)
    labels = fenics.project(fenics.Expression('(x[0]>=0.0) ? (1.0) : (2.0)', degree=1), fenics.FunctionSpace(mesh, "DG", 1))
    tissue_map = {0: 'outside', 1: 'A', 2: 'B'}
    dirichlet_bcs = {'clamped_outside': {'bc_value': fenics.Constant((0.0, 0.0)), 'named_boundary': 'boundary_all',
                                         'subspace_id': 0}}
    u_0_conc_expr = fenics.Expression('sqrt(pow(x[0]-x0,2)+pow(x[1]-y0,2)) < 0.4 ? (1.0) : (0.0)', degree=1, x0=2.5, y0=2.5)
    sim = TumorGrowth(mesh)
    sim.setup_global_parameters(label_function=labels, domain_names=tissue_map, boundaries={'boundary_all': Boundary()},
                                dirichlet_bcs=dirichlet_bcs, von_neumann_bcs={})
    sim.setup_model_parameters(iv_expression={0: fenics.Constant((0.0, 0.0)), 1: u_0_conc_expr},
                               diffusion={'outside': 0.0, 'A': 0.1, 'B': 0.0}, coupling={'outside': 0.0, 'A': 0.2, 'B': 0.0},
                               proliferation={'outside': 0.0, 'A': 0.1, 'B': 0.0}, E={'outside': 10E6, 'A': 0.001, 'B': 0.001},
                               poisson={'outside': 0.49, 'A': 0.40, 'B': 0.10}, sim_time=10, sim_time_step=1)
    sim.run(save_method=save_method, plot=True, output_dir=str(tmp_path), clear_all=True)
    return sim


def test_reference_script_runs_and_matches_oracle(tmp_path):
    sim = _script(tmp_path, "xdmf")
    form = sim.solver.problem.form
    mesh = sim.mesh
    t = form.table
    bc = sim.bcs.dirichlet_bcs[0]
    order = np.argsort(bc.dofs)
    prob = fem.Problem(mesh.coords, mesh.cells, form.cell_mat, fem.Materials(t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4]),
                       1.0, bc_dofs=bc.dofs[order], bc_vals=bc.values[order])
    x0 = sim.results.get_result(0).get_field().vector().get_local()
    recs, _ = osolver.run(prob, x0, 10, linear="lu", rtol=1e-12, atol=1e-14)
    assert sim.results.get_recording_steps() == list(range(11))
    # default tolerances (SNES rtol 1e-9 as the reference; KSP rtol 1e-10): <= 1e-6 relative L2, i.e. inside the
    # reference's own solver noise; the tight-tolerance 1e-8 parity is asserted in test_gpu_parity.py
    for k in range(1, 11):
        x = sim.results.get_result(k).get_field().vector().get_local().reshape(-1, 3)
        ref = recs[k][2].reshape(-1, 3)
        assert np.linalg.norm(x[:, 2] - ref[:, 2]) / np.linalg.norm(ref[:, 2]) < 1e-6
        assert np.linalg.norm(x[:, :2] - ref[:, :2]) / np.linalg.norm(ref[:, :2]) < 1e-6
    for f in ("solution.xdmf", "solution.h5", "solution_timeseries.h5"):
        assert os.path.exists(os.path.join(str(tmp_path), f)), f
    # reload_from_hdf5 restores the records (simulation_base.py:319-325)
    sim.reload_from_hdf5(os.path.join(str(tmp_path), "solution_timeseries.h5"))
    assert len(sim.results.get_recording_steps()) == 11


def test_rerun_with_new_parameters_reuses_the_engine(tmp_path):
    """Inverse-problem call pattern (image_based_optimization.py:531-564): many forward runs on one mesh."""
    sim = _script(tmp_path, None)
    eng = sim.solver._engine
    c_end = sim.solution.vector().get_local()[2::3].copy()
    sim.params.proliferation = 0.0          # scalar replaces the per-tissue dict
    sim.run(save_method=None, plot=False, output_dir=str(tmp_path))
    assert sim.solver._engine is eng
    assert sim.solution.vector().get_local()[2::3].sum() < c_end.sum()


def test_brain_variant_agrees_with_dict_parameters(tmp_path):
    """The reference's own cross-check (test_case_comparison_2D_atlas.py): TumorGrowth with per-tissue dicts and
    TumorGrowthBrain with per-tissue scalars give the same fields. Four tissues on a 2D strip mesh, 5 steps."""
    from glimslib_b200 import fenics_local as fenics
    from glimslib_b200.simulation.simulation_tumor_growth import TumorGrowth
    from glimslib_b200.simulation.simulation_tumor_growth_brain import TumorGrowthBrain

    class Boundary(fenics.SubDomain):
        def inside(self, x, on_boundary):
            return on_boundary

    mesh = fenics.RectangleMesh(fenics.Point(0, 0), fenics.Point(8, 4), 64, 32)
    sd = fenics.MeshFunction("size_t", mesh, 2)
    xm = mesh.cell_midpoints()[:, 0]
    sd.array()[:] = np.where(xm < 2, 1, np.where(xm < 4, 2, np.where(xm < 6, 3, 4)))
    tm = {1: "CSF", 2: "GM", 3: "WM", 4: "Ventricles"}
    bcs = {"clamped": {"bc_value": fenics.Constant((0.0, 0.0)), "named_boundary": "all", "subspace_id": 0}}
    ivs = {0: fenics.Constant((0.0, 0.0)),
           1: fenics.Expression("exp(-a*pow(x[0]-x0,2)-a*pow(x[1]-y0,2))", degree=1, a=2.0, x0=4.0, y0=2.0)}
    a = TumorGrowth(mesh)
    a.setup_global_parameters(subdomains=sd, domain_names=tm, boundaries={"all": Boundary()}, dirichlet_bcs=bcs)
    a.setup_model_parameters(iv_expression=ivs, sim_time=5, sim_time_step=1,
                             E={"CSF": 1e-3, "GM": 3e-3, "WM": 3e-3, "Ventricles": 1e-3},
                             poisson={"CSF": 0.47, "GM": 0.4, "WM": 0.4, "Ventricles": 0.3},
                             diffusion={"CSF": 0, "GM": 0.02, "WM": 0.1, "Ventricles": 0},
                             proliferation={"CSF": 0, "GM": 0.05, "WM": 0.05, "Ventricles": 0},
                             coupling={"CSF": 0.1, "GM": 0.1, "WM": 0.1, "Ventricles": 0.1})
    a.run(save_method=None, plot=False, output_dir=str(tmp_path / "a"))
    b = TumorGrowthBrain(mesh)
    b.setup_global_parameters(subdomains=sd, domain_names=tm, boundaries={"all": Boundary()}, dirichlet_bcs=bcs)
    b.setup_model_parameters(iv_expression=ivs, sim_time=5, sim_time_step=1, E_GM=3e-3, E_WM=3e-3, E_CSF=1e-3, E_VENT=1e-3,
                             nu_GM=0.4, nu_WM=0.4, nu_CSF=0.47, nu_VENT=0.3, D_GM=0.02, D_WM=0.1, rho_GM=0.05,
                             rho_WM=0.05, coupling=0.1)
    b.run(save_method=None, plot=False, output_dir=str(tmp_path / "b"))
    assert a.results.get_recording_steps() == b.results.get_recording_steps() == list(range(6))
    xa, xb = a.solution.vector().get_local(), b.solution.vector().get_local()
    assert np.abs(xa).max() > 0
    assert fenics.errornorm(a.solution, b.solution) <= 1e-9 * fenics.norm(a.solution)
    # comparison metric of the reference (helper_classes.py:2001-2013) on the concentration sub-function
    ca, cb = a.solution.sub(1, deepcopy=True), b.solution.sub(1, deepcopy=True)
    assert fenics.errornorm(ca, cb) <= 1e-9 * fenics.norm(ca)


def test_postprocess_fields_match_numpy(tmp_path):
    """N2: per-cell strain / stress / pressure / von Mises / Jacobians from the device kernel equal a numpy evaluation
    of math_linear_elasticity.py:12-46 on the same solution; nodal fields are the volume-weighted averages."""
    sim = _script(tmp_path, None)
    sim.init_postprocess(str(tmp_path / "pp"))
    pp = sim.postprocess
    mesh, form = sim.mesh, sim.solver.problem.form
    x = sim.results.get_result(10).get_field().vector().get_local().reshape(-1, 3)
    G, V = fem.geometry(mesh.coords, mesh.cells)
    gu = np.einsum("ebi,ebj->eij", x[mesh.cells][:, :, :2], G)
    eps = 0.5 * (gu + gu.transpose(0, 2, 1))
    mu, lam = form.table[form.cell_mat, 0], form.table[form.cell_mat, 1]
    sig = 2 * mu[:, None, None] * eps + (lam * np.trace(eps, axis1=1, axis2=2))[:, None, None] * np.eye(2)
    s_dev = sig - (np.trace(sig, axis1=1, axis2=2) / 3.0)[:, None, None] * np.eye(2)
    vm = np.sqrt(1.5 * (s_dev ** 2).sum(axis=(1, 2)))
    detF = np.linalg.det(np.eye(2)[None] + gu)
    got = pp.get_stress_tensor(10, cellwise=True).vector().get_local().reshape(-1, 2, 2)
    assert np.abs(got - sig).max() <= 1e-13 * np.abs(sig).max()
    assert np.abs(pp.get_van_mises_stress(10, cellwise=True).vector().get_local() - vm).max() <= 1e-13 * vm.max()
    assert np.abs(pp.get_total_jacobian(10, cellwise=True).vector().get_local() - detF).max() <= 1e-14
    p_cell = np.trace(sig, axis1=1, axis2=2) / 3.0
    lumped = np.bincount(mesh.cells.ravel(), weights=np.repeat(V, 3), minlength=mesh.num_vertices())
    p_node = np.bincount(mesh.cells.ravel(), weights=np.repeat(V * p_cell, 3), minlength=mesh.num_vertices()) / lumped
    assert np.abs(pp.get_pressure(10, lumped=True).vector().get_local() - p_node).max() <= 1e-12 * np.abs(p_node).max()
    assert pp.get_displacement_norm(10).vector().get_local().max() > 0
    # default getters = the reference's consistent-mass projections (oracle/postprocess.py), through the drop-in classes
    from oracle import postprocess as opp
    ref = opp.derived_fields(mesh.coords, mesh.cells, form.cell_mat, form.table, x.ravel())
    rel = lambda a, b: np.abs(np.asarray(a).ravel() - np.asarray(b).ravel()).max() / np.abs(b).max()
    assert rel(pp.get_stress_tensor(10).vector().get_local(), ref["stress"]) < 1e-9
    assert rel(pp.get_pressure(10).vector().get_local(), ref["pressure"]) < 1e-9
    assert rel(pp.get_van_mises_stress(10).vector().get_local(), ref["von_mises"]) < 1e-8
    assert rel(pp.get_total_jacobian(10).vector().get_local(), ref["total_jacobian"]) < 1e-9
    assert rel(pp.get_growth_induced_jacobian(10).vector().get_local(), ref["growth_jacobian"]) < 1e-8
    assert rel(pp.get_logistic_growth(10).vector().get_local(), ref["logistic_growth"]) < 1e-9
    assert rel(pp.get_displacement_norm(10).vector().get_local(), ref["displacement_norm"]) < 1e-8
    assert rel(pp.get_concentration_deformed_configuration(10).vector().get_local(), ref["concentration_deformed"]) < 1e-8


@pytest.mark.parametrize("d", [2, 3])
def test_projected_derived_fields_match_the_reference_projections(d):
    """the post-processing fields as the reference computes them -- fenics.project(expr, V), i.e. the
    consistent-mass L2 projection onto P1 (helper_classes.py:1566-1618, 1736-1786) -- against oracle/postprocess.py, which
    integrates the UFL text of math_linear_elasticity.py:12-46 literally with a degree-5 rule and solves M q = b by sparse LU:
    device-only fields (glims_project_fields) <= 1e-9, the staged ones (pressure / von Mises from the projected stress,
    growth Jacobian from the projected expansion, displacement norm; host load + glims_mass_solve) <= 1e-8."""
    from oracle import postprocess as opp, meshes
    from glimslib_b200.engine import Engine
    from glimslib_b200.backend import projection
    from glimslib_b200 import mesh as gmesh
    rng = np.random.default_rng(41)
    if d == 2:
        coords, cells = meshes.rectangle_mesh((0, 0), (1.0, 1.3), 9, 8)
    else:
        coords, cells = meshes.box_mesh((0, 0, 0), (1, 1.2, 0.8), 5, 4, 5)
    bv = meshes.boundary_vertices(cells, len(coords))
    interior = np.ones(len(coords), bool)
    interior[bv] = False
    coords = coords.copy()
    coords[interior] += 0.03 * (rng.random((interior.sum(), d)) - 0.5)
    cell_mat = rng.integers(0, 3, len(cells)).astype(np.int32)
    mats = fem.Materials.from_E_nu([3e-3, 1e-3, 2e-3], [0.45, 0.3, 0.49], [0.1, 0.02, 0.0], [0.2, 0.05, 0.0], [0.15, 0.0, 0.3])
    table = mats.table()
    nb = d + 1
    x = np.zeros((len(coords), nb))
    x[:, :d] = 0.05 * rng.standard_normal((len(coords), d))
    x[:, d] = rng.random(len(coords))
    ref = opp.derived_fields(coords, cells, cell_mat, table, x.ravel())
    eng = Engine(coords, cells, cell_mat)
    eng.set_materials(table)
    eng.set_state(x.ravel())
    f = eng.project_fields()
    rel = lambda a, b: np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)
    assert rel(f["strain"], ref["strain"]) < 1e-9
    assert rel(f["stress"], ref["stress"]) < 1e-9
    assert rel(f["total_jacobian"], ref["total_jacobian"]) < 1e-9
    assert rel(f["logistic_growth"], ref["logistic_growth"]) < 1e-9
    assert rel(np.trace(f["stress"], axis1=1, axis2=2) / 3.0, ref["pressure"]) < 1e-9
    # generic mass solve + host load: von Mises of the projected stress
    mesh = gmesh.SimplexMesh(coords, cells)
    sh, eye = f["stress"], np.eye(d)

    def vm(cc, lam, sl):
        s = np.einsum("qa,eaij->eqij", lam, sh[cc])
        dev = s - (np.trace(s, axis1=2, axis2=3) / 3.0)[:, :, None, None] * eye
        return np.sqrt(1.5 * np.einsum("eqij,eqij->eq", dev, dev))
    got = eng.mass_solve(projection.load_vector(mesh, vm))[:, 0]
    assert rel(got, ref["von_mises"]) < 1e-8
    gam = table[cell_mat, 4]
    mech = eng.mass_solve(projection.load_vector(mesh, lambda cc, lam, sl: np.einsum("qa,ea->eq", lam, x[:, d][cc]) * gam[sl][:, None]))[:, 0]
    assert rel(mech, ref["mech_expansion"]) < 1e-8
    jac = eng.mass_solve(projection.load_vector(mesh, lambda cc, lam, sl: (1.0 + np.einsum("qa,ea->eq", lam, mech[cc])) ** d))[:, 0]
    assert rel(jac, ref["growth_jacobian"]) < 1e-8
    eng.close()
