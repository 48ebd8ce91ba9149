"""CPU tests of the drop-in host layer (no GPU): the known answers the reference's own unit tests hold at this
boundary and the plumbing of setup_global_parameters / setup_model_parameters."""
import logging

import numpy as np
import pytest

from glimslib_b200 import fenics_local as fenics
from glimslib_b200.backend import cexpr
from glimslib_b200.simulation.simulation_tumor_growth import TumorGrowth
from glimslib_b200.simulation.simulation_tumor_growth_brain import TumorGrowthBrain
from glimslib_b200.simulation_helpers.helper_classes import (BoundaryConditions, FunctionSpace, Results, SubDomains,
                                                             TimeSeriesMultiData)


class Boundary(fenics.SubDomain):
    def inside(self, x, on_boundary):
        return on_boundary


class BoundaryRight(fenics.SubDomain):
    def inside(self, x, on_boundary):
        return on_boundary and x[0] >= 5.0 - 1e-10


def _labelled_mesh(nx=5, ny=5):
    mesh = fenics.RectangleMesh(fenics.Point(-5, -5), fenics.Point(5, 5), nx, ny)
    V = fenics.FunctionSpace(mesh, "DG", 1)
    labels = fenics.project(fenics.Expression("(x[0]>=0) ? (1.0) : (2.0)", degree=1), V)
    return mesh, labels


def test_cexpr_translator():
    X = np.array([[0.3, -1.0], [2.5, 2.5], [-0.1, 4.0]])
    t = cexpr.parse("sqrt(pow(x[0]-x0,2)+pow(x[1]-y0,2)) < 0.4 ? (1.0) : (0.0)")
    assert cexpr.evaluate(t, X, dict(x0=2.5, y0=2.5)).tolist() == [0.0, 1.0, 0.0]
    t = cexpr.parse("exp(-a*pow(x[0]-x0, 2) - a*pow(x[1]-y0, 2))")
    assert np.allclose(cexpr.evaluate(t, X, dict(a=0.5, x0=0, y0=0)), np.exp(-0.5 * (X ** 2).sum(axis=1)))
    assert cexpr.evaluate(cexpr.parse("x[0] > 0 && x[1] > 0 || x[0] < -0.05"), X, {}).tolist() == [0.0, 1.0, 1.0]
    with pytest.raises(cexpr.CExprError):
        cexpr.evaluate(cexpr.parse("__import__(1)"), X, {})
    with pytest.raises(cexpr.CExprError):
        cexpr.parse("x[0]; 1")


def test_subdomain_labels_known_answer():
    """test_unit_subDomains.py:36-43: labels from (x[0]>=0)?1:2 on a 5x5 mesh give the id set {1,2}."""
    mesh, labels = _labelled_mesh()
    sd = SubDomains(mesh)
    sd.setup_subdomains(label_function=labels)
    assert set(np.unique(sd.subdomains.array())) == {1, 2}


def test_interface_and_named_boundary_known_answers():
    """test_unit_subDomains.py:51-54,69-74: interface id dict values {0,1,2,3}; ids present {2,3}; the
    tissue/tumor interface has ny facets and the all-around named boundary 2(nx+ny)."""
    nx = ny = 5
    mesh, labels = _labelled_mesh(nx, ny)
    sd = SubDomains(mesh)
    sd.setup_subdomains(label_function=labels)
    sd.setup_boundaries(tissue_map={0: "outside", 1: "tissue", 2: "tumor"}, boundary_fct_dict={"boundary_1": Boundary()})
    assert set(sd.subdomain_boundaries_id_dict.values()) == {0, 1, 2, 3}
    assert set(np.unique(sd.subdomain_boundaries.array())) == {2, 3}
    assert np.sum(sd.subdomain_boundaries.array() == sd.subdomain_boundaries_id_dict["tissue_tumor"]) == ny
    assert np.sum(sd.named_boundaries.array() == sd.named_boundaries_id_dict["boundary_1"]) == 2 * (nx + ny)


def test_label_rule_puts_the_interface_one_column_left():
    """int(DG1 value at the midpoint) makes cells with a vertex on x=0 and the rest at value 2
    evaluate to 4/3 or 5/3 -> 1, so on the 50x50 case 26 of 50 columns carry label 1."""
    mesh = fenics.RectangleMesh(fenics.Point(-5, -5), fenics.Point(5, 5), 50, 50)
    labels = fenics.project(fenics.Expression("(x[0]>=0.0) ? (1.0) : (2.0)", degree=1), fenics.FunctionSpace(mesh, "DG", 1))
    sd = SubDomains(mesh)
    sd.setup_subdomains(label_function=labels)
    assert np.bincount(sd.subdomains.array()).tolist() == [0, 2600, 2400]


def test_von_neumann_terms_equal_hand_written_boundary_integrals():
    """test_unit_boundaryConditions.py:90-108: the Neumann term assembled through the BC dictionary equals the
    hand-written ds(id) integrals: sum_i g_i * |boundary_i| for constant data (7 places)."""
    mesh, labels = _labelled_mesh(10, 10)
    sim = TumorGrowth(mesh)
    sim.setup_global_parameters(label_function=labels, domain_names={0: "outside", 1: "tissue", 2: "tumor"},
                                boundaries={"all": Boundary(), "right": BoundaryRight()},
                                dirichlet_bcs={}, von_neumann_bcs={
                                    "flux_right": {"bc_value": fenics.Constant(2.0), "named_boundary": "right", "subspace_id": 1},
                                    "trac_right": {"bc_value": fenics.Constant((1.0, -3.0)), "named_boundary": "right", "subspace_id": 0}})
    sim.setup_model_parameters(iv_expression={0: fenics.Constant((0.0, 0.0)), 1: fenics.Constant(0.0)}, diffusion=0.5,
                               coupling=0.1, proliferation=0.1, E=1.0, poisson=0.3, sim_time=1, sim_time_step=0.25)
    u0 = sim.params.create_initial_value_function()
    sim._setup_problem(u0)
    f = sim.solver.problem.form.load_vector().reshape(-1, 3)
    length = 10.0          # the x = 5 edge
    # 'right' is marked after 'all' on the same facet function, so ds(right) covers the whole right edge
    assert f[:, 0].sum() == pytest.approx(1.0 * length, abs=1e-7)
    assert f[:, 1].sum() == pytest.approx(-3.0 * length, abs=1e-7)
    assert f[:, 2].sum() == pytest.approx(0.25 * 0.5 * 2.0 * length, abs=1e-7)      # dt * D * g * |boundary|


def test_dirichlet_key_handling_quirk_q1():
    """helper_classes.py:703-721: only 'boundary', 'subdomain_boundary', 'named_boundary' are recognised;
    'boundary_name' / 'boundary_id' specs are skipped silently."""
    mesh, labels = _labelled_mesh()
    sim = TumorGrowth(mesh)
    sim.setup_global_parameters(label_function=labels, domain_names={0: "outside", 1: "tissue", 2: "tumor"},
                                boundaries={"boundary_all": Boundary()},
                                dirichlet_bcs={"a": {"bc_value": fenics.Constant((0.0, 0.0)), "boundary_name": "boundary_all", "subspace_id": 0},
                                               "b": {"bc_value": fenics.Constant((0.0, 0.0)), "named_boundary": "boundary_all", "subspace_id": 0},
                                               "c": {"bc_value": fenics.Constant(0.0), "subdomain_boundary": "tissue_tumor", "subspace_id": 1},
                                               "d": {"bc_value": fenics.Constant((0.0, 0.0)), "boundary": Boundary(), "subspace_id": 0}})
    assert len(sim.bcs.dirichlet_bcs) == 3
    b, c, d = sim.bcs.dirichlet_bcs
    assert np.array_equal(np.sort(b.dofs), np.sort(d.dofs)) and len(b.dofs) == 2 * 20
    assert len(c.dofs) == 6 and np.all(c.dofs % 3 == 2)


def test_missing_required_parameter_only_warns_quirk_q7(caplog):
    mesh, labels = _labelled_mesh()
    sim = TumorGrowth(mesh)
    sim.setup_global_parameters(label_function=labels, domain_names={0: "outside", 1: "tissue", 2: "tumor"})
    with caplog.at_level(logging.WARNING):
        sim.setup_model_parameters(iv_expression={0: fenics.Constant((0.0, 0.0)), 1: fenics.Constant(0.0)},
                                   diffusion=0.1, sim_time=1, sim_time_step=1)
    assert not hasattr(sim.params, "diffusion")
    assert any("incomplete" in r.message for r in caplog.records)


def test_parameter_dicts_become_a_label_keyed_material_table():
    """Quirk Q2: the table is keyed by label id, not by dict position (unsorted tissue map)."""
    mesh = fenics.RectangleMesh(fenics.Point(0, 0), fenics.Point(4, 1), 4, 1)
    sd_fun = fenics.MeshFunction("size_t", mesh, 2)
    sd_fun.array()[:] = np.repeat([1, 3, 2, 4], 2)
    tm = {1: "CSF", 3: "WM", 2: "GM", 4: "Ventricles"}
    sim = TumorGrowth(mesh)
    sim.setup_global_parameters(subdomains=sd_fun, domain_names=tm)
    vals = {"CSF": 1.0, "WM": 3.0, "GM": 2.0, "Ventricles": 4.0}
    sim.setup_model_parameters(iv_expression={0: fenics.Constant((0.0, 0.0)), 1: fenics.Constant(0.0)}, diffusion=vals,
                               coupling=vals, proliferation=vals, E=vals, poisson={k: 0.3 for k in vals},
                               sim_time=1, sim_time_step=1)
    sim._setup_problem(sim.params.create_initial_value_function())
    form = sim.solver.problem.form
    per_cell_D = form.table[form.cell_mat, 2]
    assert per_cell_D.tolist() == np.repeat([1.0, 3.0, 2.0, 4.0], 2).tolist()


def test_brain_variant_builds_the_same_table_as_dict_parameters():
    """TumorGrowth(dict params) == TumorGrowthBrain(per-tissue scalars) (test_case_comparison_2D_atlas.py)."""
    mesh = fenics.RectangleMesh(fenics.Point(0, 0), fenics.Point(4, 1), 4, 1)
    sd_fun = fenics.MeshFunction("size_t", mesh, 2)
    sd_fun.array()[:] = np.repeat([1, 2, 3, 4], 2)
    tm = {1: "CSF", 2: "GM", 3: "WM", 4: "Ventricles"}
    ivs = {0: fenics.Constant((0.0, 0.0)), 1: fenics.Constant(0.0)}
    a = TumorGrowth(mesh)
    a.setup_global_parameters(subdomains=sd_fun, domain_names=tm)
    a.setup_model_parameters(iv_expression=ivs, sim_time=2, sim_time_step=1,
                             E={"CSF": 1e-3, "GM": 3e-3, "WM": 3e-3, "Ventricles": 1e-3},
                             poisson={"CSF": 0.47, "GM": 0.4, "WM": 0.4, "Ventricles": 0.3},
                             diffusion={"CSF": 0, "GM": 0.02, "WM": 0.1, "Ventricles": 0},
                             proliferation={"CSF": 0, "GM": 0.05, "WM": 0.05, "Ventricles": 0},
                             coupling={"CSF": 0.1, "GM": 0.1, "WM": 0.1, "Ventricles": 0.1})
    b = TumorGrowthBrain(mesh)
    b.setup_global_parameters(subdomains=sd_fun, domain_names=tm)
    b.setup_model_parameters(iv_expression=ivs, sim_time=2, sim_time_step=1, E_GM=3e-3, E_WM=3e-3, E_CSF=1e-3, E_VENT=1e-3,
                             nu_GM=0.4, nu_WM=0.4, nu_CSF=0.47, nu_VENT=0.3, D_GM=0.02, D_WM=0.1, rho_GM=0.05,
                             rho_WM=0.05, coupling=0.1)
    a._setup_problem(a.params.create_initial_value_function())
    b._setup_problem(b.params.create_initial_value_function())
    assert np.allclose(a.solver.problem.form.table, b.solver.problem.form.table, rtol=1e-15)


def test_time_series_container_round_trip(tmp_path):
    """test_unit_timeSeriesMultiData.py:95-122: save_to_hdf5 / load_from_hdf5 round trip (np.allclose)."""
    mesh, _ = _labelled_mesh()
    fs = FunctionSpace(mesh)
    cell = mesh.ufl_cell()
    fs.init_function_space(fenics.MixedElement([fenics.VectorElement("Lagrange", cell, 1), fenics.FiniteElement("Lagrange", cell, 1)]),
                           {0: "displacement", 1: "concentration"})
    ts = TimeSeriesMultiData()
    ts.register_time_series("solution", fs)
    rng = np.random.default_rng(0)
    fields = []
    for k in range(3):
        f = fenics.Function(fs.function_space)
        f.vector().set_local(rng.standard_normal(f.vector().size()))
        fields.append(f.vector().get_local())
        ts.add_observation("solution", f, time=k, time_step=k, recording_step=k)
    path = str(tmp_path / "ts.h5")
    ts.save_to_hdf5(path, replace=True)
    ts2 = TimeSeriesMultiData()
    ts2.register_time_series("solution", fs)
    ts2.load_from_hdf5(path)
    assert ts2.get_all_recording_steps("solution") == [0, 1, 2]
    for k in range(3):
        assert np.allclose(ts2.get_observation("solution", k).get_field().vector().get_local(), fields[k])
        assert ts2.get_observation("solution", k).get_time_step() == k


def test_no_gpu_fails_loudly_not_silently(tmp_path, have_gpu):
    """Without a CUDA device the run must raise (no CPU fallback), not report 'did not converge'."""
    if have_gpu:
        pytest.skip("GPU present")
    mesh, labels = _labelled_mesh()
    sim = TumorGrowth(mesh)
    sim.setup_global_parameters(label_function=labels, domain_names={0: "outside", 1: "tissue", 2: "tumor"})
    sim.setup_model_parameters(iv_expression={0: fenics.Constant((0.0, 0.0)), 1: fenics.Constant(0.1)}, diffusion=0.1,
                               coupling=0.1, proliferation=0.1, E=1.0, poisson=0.3, sim_time=1, sim_time_step=1)
    from glimslib_b200.engine import EngineError
    with pytest.raises(EngineError):
        sim.run(save_method=None, plot=False, output_dir=str(tmp_path))


def test_mesh_hdf5_round_trip(tmp_path):
    """data_io.save_mesh_hdf5 / read_mesh_hdf5 (reference data_io.py:663-713): mesh + cell and facet labels."""
    from glimslib_b200.utils import data_io as dio
    mesh, labels = _labelled_mesh(4, 3)
    sd = SubDomains(mesh)
    sd.setup_subdomains(label_function=labels)
    sd.setup_boundaries(tissue_map={0: "outside", 1: "a", 2: "b"}, boundary_fct_dict={"all": Boundary()})
    p = str(tmp_path / "mesh.h5")
    dio.save_mesh_hdf5(mesh, p, subdomains=sd.subdomains, boundaries=sd.named_boundaries)
    m2, s2, b2 = dio.read_mesh_hdf5(p)
    assert np.array_equal(m2.cells, mesh.cells) and np.allclose(m2.coords, mesh.coords)
    assert np.array_equal(s2.array(), sd.subdomains.array()) and np.array_equal(b2.array(), sd.named_boundaries.array())
    V = fenics.FunctionSpace(m2, "CG", 1)
    f = fenics.project(fenics.Expression("x[0]+2*x[1]", degree=1), V)
    dio.save_functions_hdf5({"f": f}, str(tmp_path / "f.h5"), time_step=0)
    g = dio.read_function_hdf5("f", V, str(tmp_path / "f.h5"))
    assert np.array_equal(g.vector().get_local(), f.vector().get_local())


def test_assemble_scalar_functionals_exactly():
    """fenics.assemble of coefficient products over dx / ds measures (what BoundaryConditions.implement_von_neumann_bc
    returns for a coefficient, helper_classes.py:861-908): products of up to three per-cell affine factors are exact."""
    mesh = fenics.UnitSquareMesh(6, 5)
    V = fenics.FunctionSpace(mesh, "CG", 1)
    x = fenics.project(fenics.Expression("x[0]", degree=1), V)
    y = fenics.project(fenics.Expression("x[1]", degree=1), V)
    one = fenics.Constant(1.0)
    assert abs(fenics.assemble(x * y * fenics.dx) - 0.25) < 1e-14
    assert abs(fenics.assemble(2.0 * x * x * y * fenics.dx) - 2.0 / 6.0) < 1e-14
    # exterior facets: the whole boundary, then the top edge only through a facet MeshFunction
    assert abs(fenics.assemble(one * x * fenics.ds) - 2.0) < 1e-14           # int x over the four edges: 0.5 + 0.5 + 1 + 0

    class Top(fenics.SubDomain):
        def inside(self, p, on_boundary):
            return on_boundary and p[1] > 1.0 - 1e-12
    mf = fenics.MeshFunction("size_t", mesh, 1)
    mf.set_all(0)
    Top().mark(mf, 7)
    ds = fenics.ds(subdomain_data=mf)
    assert ds.subdomain_data() is mf
    assert abs(fenics.assemble(x * y * ds(7)) - 0.5) < 1e-14
    assert abs(fenics.assemble(x * ds(7) + fenics.Constant(3.0) * y * ds(7)) - 3.5) < 1e-14
    # cells by label
    lab = fenics.MeshFunction("size_t", mesh, 2)
    lab.array()[:] = (mesh.cell_midpoints()[:, 0] > 0.5).astype(int)
    dx = fenics.dx(subdomain_data=lab)
    assert abs(fenics.assemble(one * one * dx(1)) - 0.5) < 1e-14
    with pytest.raises(TypeError):
        fenics.assemble(fenics.Constant((1.0, 2.0)) * x * fenics.dx)


def test_small_data_io_helpers(tmp_path):
    """data_io.py:132-143, 277-308, 763-800 and file_utils.py:5-21: dof maps, coordinate lookup, function + mesh files."""
    from glimslib_b200.utils import data_io as dio, file_utils as fu
    assert fu.get_file_extension("/a/b.c/file.vtu") == "vtu" and fu.get_file_extension("/a/b.c/dir") is None
    assert fu.ensure_dir_exists(str(tmp_path / "x" / "y.h5")) == str(tmp_path / "x") and (tmp_path / "x").is_dir()
    mesh, labels = _labelled_mesh(4, 3)
    el = fenics.MixedElement([fenics.VectorElement("Lagrange", mesh.ufl_cell(), 1), fenics.FiniteElement("Lagrange", mesh.ufl_cell(), 1)])
    W = fenics.FunctionSpace(mesh, el)
    by = dio.get_dofs_by_subspace(W)
    assert sorted(by) == [0, 1] and len(by[0]) == 2 * mesh.num_vertices() and len(by[1]) == mesh.num_vertices()
    assert sorted(np.concatenate([by[0], by[1]]).tolist()) == list(range(W.dim()))
    xy = dio.get_dof_coordinate_map(W)
    assert xy.shape == (W.dim(), 2)
    hit = dio.get_dofs_from_coord(xy, (5.0, 5.0))
    assert len(hit) == 3 and np.allclose(xy[hit], [5.0, 5.0])                 # u_x, u_y, c at the corner vertex
    assert dio.get_dofs_from_coord(xy, (0.123, 0.456)) is None
    V = fenics.FunctionSpace(mesh, "Lagrange", 1)
    f = fenics.project(fenics.Expression("x[0]*x[1]+1", degree=1), V)
    assert dio.get_value_dimension_from_function(f) == 1
    U = fenics.VectorFunctionSpace(mesh, "Lagrange", 1)
    assert dio.get_value_dimension_from_function(fenics.project(fenics.Constant((1.0, 2.0)), U)) == 2
    p = str(tmp_path / "out" / "conc.h5")
    dio.save_function_mesh(f, p, labelfunction=labels)
    g, m2, sd2, _ = dio.load_function_mesh(p)
    assert np.array_equal(m2.cells, mesh.cells) and np.array_equal(g.vector().get_local(), f.vector().get_local())
    assert set(np.unique(sd2.array())) == {1, 2}
