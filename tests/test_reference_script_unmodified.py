"""The reference's own example script, UNMODIFIED, exec'd from /root/reference against the drop-in package
(VERDICT r01 "missing" item 2): `glimslib` resolves to this repository's alias package, so every import, class, method and
attribute the script touches -- testing_config (`from glimslib.config import *`), fenics.set_log_level(fenics.PROGRESS),
SubDomain, RectangleMesh, Expression strings, project, TumorGrowth.setup_global_parameters / setup_model_parameters /
run(save_method='vtk', plot=True, clear_all=True), dio.merge_VTUs, init_postprocess, postprocess.plot_all -- has to exist
with the reference's meaning.  There is no GPU where the reference tree lives and no reference tree on the GPU box, so the
device engine is replaced by the oracle-backed stand-in of tests/oracle_engine.py (the product itself has no CPU path);
tests/test_gpu_dropin.py runs the same script body on the real engine."""
import os
import runpy
import sys

import numpy as np
import pytest

REF = "/root/reference"
SCRIPT = os.path.join(REF, "test_cases", "test_simulation_tumor_growth", "test_case_simulation_tumor_growth_2D_subdomains.py")
pytestmark = pytest.mark.skipif(not os.path.exists(SCRIPT), reason="reference tree not present (GPU box)")


def test_unmodified_reference_script_runs_on_the_dropin(tmp_path, monkeypatch):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    from oracle import fem, solver as osolver
    import importlib
    monkeypatch.setenv("GLIMSLIB_OUTPUT_DIR", str(tmp_path))
    import glimslib                                   # the alias package of this repository, imported FIRST
    import glimslib_b200.config as cfg
    importlib.reload(cfg)                             # pick up GLIMSLIB_OUTPUT_DIR
    import glimslib_b200.backend.problem as problem
    monkeypatch.setattr(problem, "Engine", OracleEngine)
    for m in [k for k in sys.modules if k == "test_cases" or k.startswith("test_cases.")]:
        monkeypatch.delitem(sys.modules, m)
    monkeypatch.syspath_prepend(REF)                  # for the script's `import test_cases....testing_config`
    import glimslib.config
    assert "glimslib_b200" in os.path.realpath(sys.modules["glimslib.config"].__file__)
    OracleEngine.instances.clear()
    g = runpy.run_path(SCRIPT, run_name="__main__")
    sim = g["sim"]
    assert type(sim).__module__ == "glimslib_b200.simulation.simulation_tumor_growth"
    out = os.path.join(str(tmp_path), "test_cases", "simulation_tumor_growth", "test_case_simulation_tumor_growth_2D_subdomains")
    assert g["output_path"] == out
    # eleven records (t = 0..10), merged VTUs per time step with both fields on the label-map mesh
    assert sim.results.get_recording_steps() == list(range(11))
    from glimslib_b200.backend import vtu
    for step in (0, 5, 10):
        m = vtu.read_vtu(os.path.join(out, "merged", "all_%05d000000.vtu" % step))
        assert set(m.point_data) >= {"concentration", "displacement"}
        assert len(m.cells["triangle"]) == 5000
    assert os.path.exists(os.path.join(out, "solution_timeseries.h5"))
    # the fields the script produced == an oracle run on the same labels / IC / BCs (script-independent set-up)
    eng = OracleEngine.instances[-1]
    t = eng.table
    prob = fem.Problem(eng.coords, eng.cells, eng.cell_mat, fem.Materials(t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4]), 1.0,
                       bc_dofs=eng.bc_dofs, bc_vals=eng.bc_vals)
    x0 = sim.results.get_result(0).get_field().vector().get_local()
    recs, _ = osolver.run(prob, x0, 10, linear="lu")
    for k in (1, 10):
        x = sim.results.get_result(k).get_field().vector().get_local()
        assert np.linalg.norm(x - recs[k][2]) <= 1e-12 * np.linalg.norm(recs[k][2])
    # labels by the helper_classes.py:441-442 rule: tissue A right of x = 0 except the column of cells touching x = 0
    lab = np.asarray(sim.subdomains.subdomains.array())
    assert set(np.unique(lab)) == {1, 2}


def _run_script(name, tmp_path, monkeypatch):
    """Exec one unmodified script of /root/reference/test_cases/test_simulation_tumor_growth on the drop-in (oracle-backed
    stand-in engine, see the module docstring); returns (script globals, the engine the run created or None)."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    import importlib
    monkeypatch.setenv("GLIMSLIB_OUTPUT_DIR", str(tmp_path))
    import glimslib                                   # noqa: F401
    import glimslib_b200.config as cfg
    importlib.reload(cfg)
    import glimslib_b200.backend.problem as problem
    monkeypatch.setattr(problem, "Engine", OracleEngine)
    for m in [k for k in sys.modules if k == "test_cases" or k.startswith("test_cases.")]:
        monkeypatch.delitem(sys.modules, m)
    monkeypatch.syspath_prepend(REF)
    OracleEngine.instances.clear()
    g = runpy.run_path(os.path.join(REF, "test_cases", "test_simulation_tumor_growth", name), run_name="__main__")
    return g, (OracleEngine.instances[-1] if OracleEngine.instances else None)


def _oracle_records(eng, x0, n_steps):
    from oracle import fem, solver as osolver
    t = eng.table
    prob = fem.Problem(eng.coords, eng.cells, eng.cell_mat, fem.Materials(t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4]), 1.0,
                       bc_dofs=eng.bc_dofs, bc_vals=eng.bc_vals)
    return osolver.run(prob, x0, n_steps, linear="lu")[0]


def test_unmodified_uniform_script(tmp_path, monkeypatch):
    """test_case_simulation_tumor_growth_2D_uniform.py: uniform parameters, named-boundary Dirichlet condition, five steps, VTK
    output, merge_VTUs(remove=True), init_postprocess + plot_all."""
    g, eng = _run_script("test_case_simulation_tumor_growth_2D_uniform.py", tmp_path, monkeypatch)
    sim = g["sim"]
    assert sim.results.get_recording_steps() == list(range(6))
    x0 = sim.results.get_result(0).get_field().vector().get_local()
    recs = _oracle_records(eng, x0, 5)
    for k in (1, 5):
        x = sim.results.get_result(k).get_field().vector().get_local()
        assert np.linalg.norm(x - recs[k][2]) <= 1e-12 * np.linalg.norm(recs[k][2])
    out = g["output_path"]
    from glimslib_b200.backend import vtu
    m = vtu.read_vtu(os.path.join(out, "merged", "all_%05d000000.vtu" % 5))
    assert set(m.point_data) >= {"concentration", "displacement"} and len(m.cells["triangle"]) == 5000
    assert not os.path.exists(os.path.join(out, "concentration", "concentration_00005000000.vtu"))      # remove=True


def test_unmodified_run_then_reload_scripts(tmp_path, monkeypatch):
    """test_case_..._2D_uniform_mpi.py writes solution_timeseries.h5 (20 steps, no VTK); test_case_..._2D_uniform_reload.py
    builds a fresh simulation, reload_from_hdf5()s that file and re-saves steps 1..19 through postprocess.save_all
    (simulation_base.py:319-325, helper_classes.py:1922-1941)."""
    g1, eng = _run_script("test_case_simulation_tumor_growth_2D_uniform_mpi.py", tmp_path, monkeypatch)
    sim1 = g1["sim"]
    assert sim1.results.get_recording_steps() == list(range(21))
    x0 = sim1.results.get_result(0).get_field().vector().get_local()
    recs = _oracle_records(eng, x0, 20)
    x20 = sim1.results.get_result(20).get_field().vector().get_local()
    assert np.linalg.norm(x20 - recs[20][2]) <= 1e-12 * np.linalg.norm(recs[20][2])
    g2, eng2 = _run_script("test_case_simulation_tumor_growth_2D_uniform_reload.py", tmp_path, monkeypatch)
    sim2 = g2["sim"]
    assert eng2 is None                                   # nothing was solved: the results come from the file
    assert sim2.results.get_recording_steps() == list(range(21))
    for k in (0, 7, 20):
        a = sim1.results.get_result(k).get_field().vector().get_local()
        b = sim2.results.get_result(k).get_field().vector().get_local()
        assert np.array_equal(a, b)
    out = g2["output_path"]
    from glimslib_b200.backend import vtu
    for k in (1, 19):
        m = vtu.read_vtu(os.path.join(out, "merged", "all_%05d000000.vtu" % k))
        assert set(m.point_data) >= {"concentration", "displacement"}
    assert not os.path.exists(os.path.join(out, "merged", "all_%05d000000.vtu" % 20))        # selection = slice(1, -1, 1)
