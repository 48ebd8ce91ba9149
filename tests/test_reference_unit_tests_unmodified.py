"""The reference's OWN unit tests of the plug-in boundary, unmodified, run against the drop-in package: the files are loaded
from /root/reference by path (their `glimslib` imports resolve to this repository's alias package), their unittest.TestCase
classes are collected and run.  SURVEY.md section 4 lists them as the reference's test strategy for the helper classes the hot
path is reached through (FunctionSpace, SubDomains, SubSpaces, BoundaryConditions, Parameters, Results, TimeSeries*).

Two tests of simulation/test_baseImplementation.py cannot pass against the reference itself and are expected to fail here in
exactly the same way:
  * test_setup_global_parameters expects 3 Dirichlet conditions, but two of its three specifications use the keys
    'boundary_id' / 'boundary_name', which BoundaryConditions._construct_dirichlet_bc (helper_classes.py:701-722) does not
    know ("Dirichlet BC incomplete -- skipping"): the reference builds 1, and so does the drop-in;
  * test_setup_model_parameters reads self.params, which its setUp never defines (AttributeError inside the test).
utils/test_unit_data_io.py needs SimpleITK (absent here) and tests image <-> function conversion, out of scope.
The reference tree does not exist on the GPU box: skipped there."""
import importlib
import importlib.util
import io
import os
import sys
import unittest

import pytest

REF = "/root/reference/glimslib"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")

FILES = {
    "simulation_helpers/test_unit_boundaryConditions.py": 3,
    "simulation_helpers/test_unit_functionSpace.py": 3,
    "simulation_helpers/test_unit_results.py": 4,
    "simulation_helpers/test_unit_simulationParameters.py": 6,
    "simulation_helpers/test_unit_subDomains.py": 5,
    "simulation_helpers/test_unit_subSpaces.py": 12,
    "simulation_helpers/test_unit_timeSeriesData.py": 4,
    "simulation_helpers/test_unit_timeSeriesDataTimePoint.py": 5,
    "simulation_helpers/test_unit_timeSeriesMultiData.py": 8,
    "simulation/test_baseImplementation.py": 2,
}
STALE_IN_THE_REFERENCE = {
    "test_setup_global_parameters": "1 != 3",
    "test_setup_model_parameters": "has no attribute 'params'",
}


def _run(rel):
    import glimslib                                        # noqa: F401  (the alias package, before anything else)
    path = os.path.join(REF, rel)
    spec = importlib.util.spec_from_file_location("refunit_" + os.path.basename(rel)[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for name, m in list(sys.modules.items()):               # every glimslib module in play is this repository's
        if (name == "glimslib" or name.startswith("glimslib.")) and getattr(m, "__file__", None):
            assert "/root/reference" not in os.path.realpath(m.__file__), name
    suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
    res = unittest.TextTestRunner(stream=io.StringIO(), verbosity=0).run(suite)
    return res


@pytest.mark.parametrize("rel", sorted(FILES))
def test_reference_unit_test_file(rel, tmp_path, monkeypatch):
    monkeypatch.setenv("GLIMSLIB_OUTPUT_DIR", str(tmp_path))
    import glimslib_b200.config as cfg
    importlib.reload(cfg)
    res = _run(rel)
    assert res.testsRun == FILES[rel]
    bad = {t.id().split(".")[-1]: tb for t, tb in res.failures + res.errors}
    for name, needle in STALE_IN_THE_REFERENCE.items():
        if rel.endswith("test_baseImplementation.py"):
            assert name in bad and needle in bad[name], (name, bad.get(name))
            bad.pop(name)
    assert not bad, "\n".join("%s: %s" % (k, v.strip().splitlines()[-1]) for k, v in bad.items())
