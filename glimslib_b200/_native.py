"""ctypes binding of ``libglims_b200.so`` (the C ABI declared in ``include/glims_b200.h``).

This is the only bridge between the Python host side and the CUDA kernels.  There is no CPU
fallback: if the shared library is missing or fails to load, importing the engine raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libglims_b200.so")

# every symbol include/glims_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "glims_default_opts", "glims_create", "glims_destroy", "glims_last_error", "glims_set_materials",
    "glims_set_dt", "glims_set_dirichlet", "glims_set_load", "glims_set_state", "glims_get_state",
    "glims_set_prev", "glims_get_prev", "glims_ndof", "glims_nnzb", "glims_nslots", "glims_state_dev",
    "glims_stream", "glims_step", "glims_assemble", "glims_get_residual", "glims_export_pattern",
    "glims_export_values", "glims_spmv", "glims_time_kernel", "glims_launch_count", "glims_cell_fields",
    "glims_nccl_unique_id", "glims_comm_init", "glims_set_halo", "glims_tile_info", "glims_tile_config", "glims_set_p2p", "glims_comm_bench",
    "glims_prepare", "glims_reset_history", "glims_project_fields", "glims_mass_solve", "glims_adjoint", "glims_set_dof_permutation", "glims_get_dof_permutation",
]

OK, ERR_ARG, ERR_CUDA, ERR_NOT_CONVERGED, ERR_STATE, ERR_NCCL = 0, -1, -2, -3, -4, -5
ASM_RESIDUAL, ASM_KCONST, ASM_KCC, ASM_JACOBIAN, ASM_ALL = 1, 2, 4, 6, 7
SOLVER_BLOCK_TRI, SOLVER_MONO_GMRES = 0, 1
PC_JACOBI, PC_AMG, PC_AMG_FP64 = 0, 1, 2
ASMK_ATOMIC, ASMK_GATHER, ASMK_SLICE, ASMK_TILE, ASMK_ROWS = 0, 1, 2, 3, 4


class SolverOpts(C.Structure):
    _fields_ = [("snes_rtol", C.c_double), ("snes_atol", C.c_double), ("snes_stol", C.c_double),
                ("max_newton", C.c_int32), ("ksp_rtol", C.c_double), ("ksp_atol", C.c_double),
                ("max_krylov", C.c_int32), ("solver", C.c_int32), ("pc", C.c_int32),
                ("asm_kernel", C.c_int32), ("lag_mechanics", C.c_int32), ("recycle", C.c_int32), ("extrapolate", C.c_int32)]


class StepStats(C.Structure):
    _fields_ = [("newton_its", C.c_int32), ("krylov_its_c", C.c_int32), ("krylov_its_u", C.c_int32),
                ("krylov_its_mono", C.c_int32), ("converged", C.c_int32), ("fnorm0", C.c_double),
                ("fnorm", C.c_double), ("ms_total", C.c_float), ("ms_assembly", C.c_float),
                ("ms_krylov", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class NativeLibraryMissing(ImportError):
    pass


_lib = None


def load():
    """Load the library once; raise loudly if it is not built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            "glimslib_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C glimslib_b200/csrc`; there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    p, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    sig = {
        "glims_default_opts": (None, [C.POINTER(SolverOpts)]),
        "glims_create": (i32, [C.POINTER(p), i32, i64, dp, i64, ip, ip, i64, i32]),
        "glims_destroy": (i32, [p]),
        "glims_last_error": (C.c_char_p, [p]),
        "glims_set_materials": (i32, [p, i32, dp]),
        "glims_set_dt": (i32, [p, dbl]),
        "glims_set_dirichlet": (i32, [p, i64, lp, dp]),
        "glims_set_load": (i32, [p, dp]),
        "glims_set_state": (i32, [p, dp]), "glims_get_state": (i32, [p, dp]),
        "glims_set_prev": (i32, [p, dp]), "glims_get_prev": (i32, [p, dp]),
        "glims_ndof": (i64, [p]), "glims_nnzb": (i64, [p]), "glims_nslots": (i64, [p]),
        "glims_state_dev": (p, [p]), "glims_stream": (p, [p]),
        "glims_step": (i32, [p, i32, C.POINTER(SolverOpts), C.POINTER(StepStats)]),
        "glims_assemble": (i32, [p, i32, i32, i32]),
        "glims_get_residual": (i32, [p, dp]),
        "glims_export_pattern": (i32, [p, lp, ip]),
        "glims_export_values": (i32, [p, dp, dp, dp]),
        "glims_spmv": (i32, [p, i32, dp, dp]),
        "glims_time_kernel": (i32, [p, i32, i32, i32, i32, C.POINTER(C.c_float)]),
        "glims_launch_count": (i64, [p]),
        "glims_cell_fields": (i32, [p, dp, dp]),
        "glims_nccl_unique_id": (i32, [p]),
        "glims_comm_init": (i32, [p, i32, i32, p]),
        "glims_set_halo": (i32, [p, i32, ip, lp, ip, lp]),
        "glims_tile_info": (i32, [p, lp]),
        "glims_tile_config": (i32, [p, i32, i32]),
        "glims_set_p2p": (i32, [p, i32]),
        "glims_comm_bench": (i32, [p, i32, i32, C.POINTER(C.c_float)]),
        "glims_prepare": (i32, [p, C.POINTER(SolverOpts)]),
        "glims_project_fields": (i32, [p, dp]),
        "glims_mass_solve": (i32, [p, i32, dp, dp]),
        "glims_adjoint": (i32, [p, i32, C.POINTER(SolverOpts), i32, dp, dp, dp, C.POINTER(C.c_double), dp]),
        "glims_reset_history": (i32, [p]),
        "glims_set_dof_permutation": (i32, [p, lp]),
        "glims_get_dof_permutation": (i32, [p, lp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    _lib = lib
    return lib


def as_dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def as_ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def as_lp(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
