"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports exactly the
symbols include/glims_b200.h declares.  No compute call is made (no GPU here)."""
import ctypes
import os
import re

from glimslib_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "glims_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(glims_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(N.SYMBOLS)


def test_library_loads_and_exports_every_symbol():
    assert os.path.exists(N.LIB_PATH), "libglims_b200.so is not built (run __graft_entry__.build())"
    lib = N.load()
    for s in header_symbols():
        assert hasattr(lib, s), s


def test_default_opts_roundtrip():
    lib = N.load()
    o = N.SolverOpts()
    lib.glims_default_opts(ctypes.byref(o))
    # DOLFIN PETScSNESSolver defaults left untouched by simulation_tumor_growth.py:126-130 [MEM]
    assert o.snes_rtol == 1e-9 and o.snes_atol == 1e-10 and o.max_newton == 50
    assert o.solver == N.SOLVER_BLOCK_TRI and o.lag_mechanics == 1


def test_null_context_is_an_error_not_a_crash():
    lib = N.load()
    assert lib.glims_set_dt(None, 1.0) == N.ERR_ARG
    assert lib.glims_destroy(None) == N.ERR_ARG
