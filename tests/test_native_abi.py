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
    # entry points added with the tile-assembly kernel and the peer-memory transport
    assert lib.glims_tile_config(None, 0, 0) == N.ERR_ARG
    assert lib.glims_tile_info(None, None) == N.ERR_ARG
    assert lib.glims_set_p2p(None, 1) == N.ERR_ARG
    assert lib.glims_comm_bench(None, 0, 1, None) == N.ERR_ARG


def test_create_rejects_bad_arguments_before_touching_cuda():
    """Argument validation happens before any CUDA call, so it is testable without a GPU."""
    import numpy as np
    lib = N.load()
    h = ctypes.c_void_p()
    xy = np.zeros((3, 2))
    cells = np.array([[0, 1, 2]], dtype=np.int32)
    mat = np.zeros(1, dtype=np.int32)
    args = (N.as_dp(xy), 1, N.as_ip(cells), N.as_ip(mat), -1, 0)
    assert lib.glims_create(ctypes.byref(h), 4, 3, *args) == N.ERR_ARG          # dim must be 2 or 3
    assert lib.glims_create(ctypes.byref(h), 2, 0, *args) == N.ERR_ARG          # empty vertex set
    assert lib.glims_create(ctypes.byref(h), 2, 3, N.as_dp(xy), 0, N.as_ip(cells), N.as_ip(mat), -1, 0) == N.ERR_ARG
    assert lib.glims_create(ctypes.byref(h), 2, 3, None, 1, N.as_ip(cells), N.as_ip(mat), -1, 0) == N.ERR_ARG
    assert lib.glims_create(None, 2, 3, *args) == N.ERR_ARG


def test_engine_validates_host_arrays_before_the_abi():
    import numpy as np
    import pytest
    from glimslib_b200.engine import Engine
    xy = np.zeros((3, 2))
    with pytest.raises(ValueError):
        Engine(xy, np.array([[0, 1, 3]]), np.zeros(1))                            # vertex index out of range
    with pytest.raises(ValueError):
        Engine(xy, np.array([[0, 1, 2, 2]]), np.zeros(1))                         # tets on 2D coordinates
    with pytest.raises(ValueError):
        Engine(np.zeros((3, 4)), np.array([[0, 1, 2]]), np.zeros(1))              # 4D coordinates
    with pytest.raises(ValueError):
        Engine(xy, np.array([[0, 1, 2]]), np.zeros(2))                            # ragged labels
