"""Backend seam -- the drop-in replacement for ``glimslib/fenics_local.py:3-25``.

The reference does ``from dolfin import *`` here and every hot-path module reaches FEniCS only through
this module.  This version re-exports the B200 backend instead: host-side problem description objects
(:mod:`glimslib_b200.backend.core`), result writers (:mod:`glimslib_b200.backend.io`) and the solver pair
``NonlinearVariationalProblem`` / ``NonlinearVariationalSolver`` whose ``solve()`` runs on the GPU.
"""
from glimslib_b200.backend.core import *            # noqa: F401,F403
from glimslib_b200.backend.core import __version__  # noqa: F401
from glimslib_b200.backend.io import HDF5File, XDMFFile, File  # noqa: F401
from glimslib_b200.backend.problem import (NonlinearVariationalProblem, NonlinearVariationalSolver,  # noqa: F401
                                            CoupledRDMechanicsForm, SolverNotConverged)
from glimslib_b200 import config

import types as _types

# dolfin's module paths of the Function class (2018+: dolfin.function.function.Function, before: dolfin.functions.Function);
# the reference's unit tests compare types against them (simulation_helpers/test_unit_subSpaces.py:116-129)
function = _types.SimpleNamespace(function=_types.SimpleNamespace(Function=Function))      # noqa: F405
functions = _types.SimpleNamespace(Function=Function)                                      # noqa: F405

if config.USE_ADJOINT:
    raise ImportError("USE_ADJOINT: the dolfin-adjoint tape is not part of the B200 hot path")


def is_version(comparison_str):
    """Same contract as fenics_local.py:12-25: '<2018.1.x', '=2017.2.x', '>2017.2.x' compare year (and major)."""
    comp, target = comparison_str[0], comparison_str[1:]
    t_year, t_major, _ = target.split(".")
    year, major, _ = __version__.split(".")
    if comp == "=":
        return year == t_year and major == t_major
    if comp == ">":
        return int(year) > int(t_year)
    if comp == "<":
        return int(year) < int(t_year)
    raise ValueError("comparison must start with one of = < >")
