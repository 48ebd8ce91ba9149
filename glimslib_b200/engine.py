"""Python handle on one ``glims_ctx`` (one GPU): the object the drop-in simulation classes drive.

It owns no numerics -- every method is a C-ABI call into the CUDA library; numpy arrays are only
the host buffers the ABI reads from / writes to.
"""
import ctypes as C

import numpy as np

from . import _native as N


class SolverNotConverged(RuntimeError):
    """Raised when the device Newton/Krylov solve fails -- ``simulation_base.py:301-305`` catches it."""


class EngineError(RuntimeError):
    pass


class Engine:
    def __init__(self, coords, cells, cell_mat, device=0, n_owned=-1, reorder=None):
        """``reorder='morton'`` (one GPU only): the library renumbers vertices and cells for locality before building its
        sparsity pattern (``mesh.locality_order``) and installs the matching dof permutation, so that every host vector,
        Dirichlet dof and derived field of this handle stays in the CALLER's numbering."""
        self._lib = N.load()
        coords = N.f64(coords)
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        cell_mat = np.ascontiguousarray(cell_mat, dtype=np.int32)
        self._new_of_old = self._cell_order = None
        if reorder not in (None, "none", "morton"):
            raise ValueError("reorder must be None or 'morton'")
        if reorder == "morton":
            if n_owned >= 0:
                raise ValueError("reorder applies to an unpartitioned mesh (reorder before partitioning instead)")
            from . import mesh as _mesh
            self._new_of_old, self._cell_order = _mesh.locality_order(coords, cells)
            old_of_new = np.empty_like(self._new_of_old)
            old_of_new[self._new_of_old] = np.arange(len(coords))
            coords = np.ascontiguousarray(coords[old_of_new])
            cells = np.ascontiguousarray(self._new_of_old[cells][self._cell_order].astype(np.int32))
            cell_mat = np.ascontiguousarray(cell_mat[self._cell_order])
        if coords.ndim != 2 or coords.shape[1] not in (2, 3):
            raise ValueError("coords must be (n_vertices, 2|3)")
        if cells.ndim != 2 or cells.shape[1] != coords.shape[1] + 1 or len(cell_mat) != len(cells):
            raise ValueError("cells must be (n_cells, dim+1) with one material index per cell")
        if cells.min() < 0 or cells.max() >= len(coords):
            raise ValueError("cell vertex index out of range")
        self.dim = coords.shape[1]
        self.nb = self.dim + 1
        self.n_vertices = len(coords)
        self.n_cells = len(cells)
        self.n_owned = self.n_vertices if n_owned < 0 else int(n_owned)
        self._h = C.c_void_p()
        rc = self._lib.glims_create(C.byref(self._h), self.dim, self.n_vertices, N.as_dp(coords), self.n_cells,
                                    N.as_ip(cells), N.as_ip(cell_mat), int(n_owned), int(device))
        if rc != 0:
            msg = self._lib.glims_last_error(self._h).decode() if self._h else "glims_create failed"
            raise EngineError("glims_create: rc=%d %s" % (rc, msg))
        self.ndof = int(self._lib.glims_ndof(self._h))
        self.opts = N.SolverOpts()
        self._lib.glims_default_opts(C.byref(self.opts))
        if self._new_of_old is not None:
            self.set_dof_permutation((self._new_of_old[:, None] * self.nb + np.arange(self.nb)[None, :]).ravel())

    # -- plumbing -----------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc == 0:
            return
        msg = self._lib.glims_last_error(self._h).decode()
        if rc == N.ERR_NOT_CONVERGED:
            raise SolverNotConverged("%s: %s" % (what, msg))
        raise EngineError("%s: rc=%d %s" % (what, rc, msg))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.glims_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- problem definition -------------------------------------------------------------------
    def set_materials(self, table):
        t = N.f64(table)
        if t.ndim != 2 or t.shape[1] != 5:
            raise ValueError("material table must be (n_mat, 5): mu, lambda, D, rho, gamma")
        self._check(self._lib.glims_set_materials(self._h, len(t), N.as_dp(t)), "set_materials")
        self._n_mat = len(t)

    def set_dt(self, dt):
        self._check(self._lib.glims_set_dt(self._h, float(dt)), "set_dt")

    def set_dirichlet(self, dofs, vals):
        d = np.ascontiguousarray(dofs, dtype=np.int64)
        v = N.f64(vals)
        if len(d) != len(v):
            raise ValueError("dofs and vals differ in length")
        self._check(self._lib.glims_set_dirichlet(self._h, len(d), N.as_lp(d), N.as_dp(v)), "set_dirichlet")

    def set_load(self, f_ext):
        if f_ext is None:
            self._check(self._lib.glims_set_load(self._h, None), "set_load")
        else:
            f = N.f64(f_ext)
            assert f.size == self.ndof
            self._check(self._lib.glims_set_load(self._h, N.as_dp(f)), "set_load")

    # -- state --------------------------------------------------------------------------------
    def _vec(self, fn, name, x=None):
        if x is None:
            out = np.empty(self.ndof)
            self._check(fn(self._h, N.as_dp(out)), name)
            return out
        a = N.f64(x).ravel()
        assert a.size == self.ndof, "%s: expected %d values" % (name, self.ndof)
        self._check(fn(self._h, N.as_dp(a)), name)

    def set_state(self, x):
        self._vec(self._lib.glims_set_state, "set_state", x)

    def get_state(self, out=None):
        """Device -> host copy of the iterate; ``out`` (C-contiguous float64, ndof) is filled in place if given."""
        if out is not None:
            assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == self.ndof
            self._check(self._lib.glims_get_state(self._h, N.as_dp(out)), "get_state")
            return out
        return self._vec(self._lib.glims_get_state, "get_state")

    def set_prev(self, x):
        self._vec(self._lib.glims_set_prev, "set_prev", x)

    def get_prev(self):
        return self._vec(self._lib.glims_get_prev, "get_prev")

    def set_dof_permutation(self, perm):
        """Caller dof numbering: ``perm[i]`` = vertex-blocked index ``v*(dim+1)+k`` of the caller's dof ``i`` (e.g. from
        DOLFIN's ``vertex_to_dof_map``); ``None`` restores the identity.  Call before :meth:`set_dirichlet`."""
        if perm is None:
            self._check(self._lib.glims_set_dof_permutation(self._h, None), "set_dof_permutation")
            return
        a = np.ascontiguousarray(perm, dtype=np.int64)
        assert a.size == self.ndof
        self._check(self._lib.glims_set_dof_permutation(self._h, N.as_lp(a)), "set_dof_permutation")

    def get_dof_permutation(self):
        out = np.empty(self.ndof, dtype=np.int64)
        self._check(self._lib.glims_get_dof_permutation(self._h, N.as_lp(out)), "get_dof_permutation")
        return out

    # -- hot path -----------------------------------------------------------------------------
    def prepare(self, **opts):
        """One-time work of the first step (K_uu / K_uc, Dirichlet elimination, AMG hierarchy, row-walk maps)."""
        for k, v in opts.items():
            setattr(self.opts, k, v)
        self._check(self._lib.glims_prepare(self._h, C.byref(self.opts)), "prepare")

    def reset_history(self):
        """Drop the projection basis / extrapolation history (matrices, hierarchy and graphs stay)."""
        self._check(self._lib.glims_reset_history(self._h), "reset_history")

    def step(self, n_steps=1, **opts):
        """``n_steps`` x (Newton-Krylov solve + ``u_previous.assign``). Returns per-step stats dicts."""
        for k, v in opts.items():
            setattr(self.opts, k, v)
        stats = (N.StepStats * n_steps)()
        rc = self._lib.glims_step(self._h, n_steps, C.byref(self.opts), stats)
        out = [s.as_dict() for s in stats]
        self.last_stats = out
        self._check(rc, "step")
        return out

    # -- building blocks ----------------------------------------------------------------------
    def assemble(self, what=N.ASM_ALL, kernel=N.ASMK_ATOMIC, apply_bc=0):
        self._check(self._lib.glims_assemble(self._h, what, kernel, apply_bc), "assemble")

    def residual(self):
        return self._vec(self._lib.glims_get_residual, "get_residual")

    def export_blocks(self):
        """(rowptr, colidx, Kuu[nnzb,d,d], Kuc[nnzb,d], Kcc[nnzb]) in CSR block order."""
        nnzb = int(self._lib.glims_nnzb(self._h))
        d = self.dim
        rowptr = np.empty(self.n_owned + 1, dtype=np.int64)
        col = np.empty(nnzb, dtype=np.int32)
        self._check(self._lib.glims_export_pattern(self._h, N.as_lp(rowptr), N.as_ip(col)), "export_pattern")
        Kuu, Kuc, Kcc = np.empty((nnzb, d, d)), np.empty((nnzb, d)), np.empty(nnzb)
        self._check(self._lib.glims_export_values(self._h, N.as_dp(Kuu), N.as_dp(Kuc), N.as_dp(Kcc)), "export_values")
        return rowptr, col, Kuu, Kuc, Kcc

    def export_jacobian(self):
        """Monolithic Jacobian as a scipy CSR matrix in vertex-blocked dof order."""
        import scipy.sparse as sp
        rowptr, col, Kuu, Kuc, Kcc = self.export_blocks()
        d, nb = self.dim, self.nb
        blocks = np.zeros((len(col), nb, nb))
        blocks[:, :d, :d] = Kuu
        blocks[:, :d, d] = Kuc
        blocks[:, d, d] = Kcc
        J = sp.bsr_matrix((blocks, col, rowptr), shape=(self.n_owned * nb, self.n_vertices * nb))
        return J.tocsr()

    def spmv(self, which, x):
        bs = {0: self.nb, 1: self.dim, 2: 1}[which]
        a = N.f64(x).ravel()
        assert a.size == self.n_vertices * bs
        y = np.empty(self.n_owned * bs)
        self._check(self._lib.glims_spmv(self._h, which, N.as_dp(a), N.as_dp(y)), "spmv")
        return y

    FIELD_NAMES = ("strain", "stress", "pressure", "von_mises", "total_jacobian", "growth_jacobian", "logistic_growth")

    def cell_fields(self, vertex=False):
        """Derived fields of the current device state. Returns a dict of per-cell arrays (or volume-weighted nodal
        averages with ``vertex=True``): strain/stress [n, d, d], the rest [n]."""
        d = self.dim
        nf = 2 * d * d + 5
        n = self.n_vertices if vertex else self.n_cells
        out = np.empty((n, nf))
        args = (None, N.as_dp(out)) if vertex else (N.as_dp(out), None)
        self._check(self._lib.glims_cell_fields(self._h, *args), "cell_fields")
        if self._new_of_old is not None:          # back to the caller's vertex / cell order
            if vertex:
                out = out[self._new_of_old]
            else:
                inv = np.empty_like(self._cell_order)
                inv[self._cell_order] = np.arange(len(inv))
                out = out[inv]
        res = {"strain": out[:, :d * d].reshape(n, d, d), "stress": out[:, d * d:2 * d * d].reshape(n, d, d)}
        for k, name in enumerate(self.FIELD_NAMES[2:]):
            res[name] = out[:, 2 * d * d + k]
        return res

    def project_fields(self):
        """The derived fields of :meth:`cell_fields` L2-projected onto P1 with the consistent mass matrix (what the
        reference's ``fenics.project`` returns), all on the device.  Dict of per-vertex arrays in the caller's order."""
        d = self.dim
        nf = 2 * d * d + 5
        n = self.n_vertices
        out = np.empty((n, nf))
        self._check(self._lib.glims_project_fields(self._h, N.as_dp(out)), "project_fields")
        if self._new_of_old is not None:
            out = out[self._new_of_old]
        res = {"strain": out[:, :d * d].reshape(n, d, d), "stress": out[:, d * d:2 * d * d].reshape(n, d, d)}
        for k, name in enumerate(self.FIELD_NAMES[2:]):
            res[name] = out[:, 2 * d * d + k]
        return res

    def mass_solve(self, load):
        """``M^-1 load`` column by column (``load``: [n_vertices, nf] in the caller's vertex order), M = P1 mass matrix."""
        a = N.f64(load)
        a = a.reshape(self.n_vertices, -1)
        if self._new_of_old is not None:
            old_of_new = np.empty_like(self._new_of_old)
            old_of_new[self._new_of_old] = np.arange(self.n_vertices)
            a = np.ascontiguousarray(a[old_of_new])
        out = np.empty_like(a)
        self._check(self._lib.glims_mass_solve(self._h, a.shape[1], N.as_dp(a), N.as_dp(out)), "mass_solve")
        return out[self._new_of_old] if self._new_of_old is not None else out

    def adjoint_gradient(self, n_steps, levels=(), level_targets=None, u_target=None, **opts):
        """Forward ``n_steps`` from the current (prev, state) and the discrete-adjoint gradient of the final-state misfit
        (``glims_adjoint``).  Returns ``(J, grad)`` with ``grad[m] = (dJ/dD_m, dJ/drho_m, dJ/dgamma_m)`` per material row."""
        for k, v in opts.items():
            setattr(self.opts, k, v)
        if self._new_of_old is not None:
            raise EngineError("adjoint_gradient: not available on a reordered engine")
        lev = N.f64(np.atleast_1d(np.asarray(levels, dtype=np.float64)))
        nl = len(lev) if len(levels) else 0
        tg = N.f64(level_targets).reshape(nl, self.n_vertices) if nl else np.zeros((1, 1))
        ut = None if u_target is None else N.f64(u_target).reshape(self.n_vertices, self.dim)
        n_mat = int(self._n_mat)
        J = C.c_double()
        grad = np.zeros((n_mat, 3))
        rc = self._lib.glims_adjoint(self._h, int(n_steps), C.byref(self.opts), nl, N.as_dp(lev) if nl else None,
                                     N.as_dp(tg) if nl else None, N.as_dp(ut) if ut is not None else None, C.byref(J), N.as_dp(grad))
        self._check(rc, "adjoint")
        return float(J.value), grad

    def time_kernel(self, kernel, variant=0, reps=10, flush_l2=True):
        ms = C.c_float()
        self._check(self._lib.glims_time_kernel(self._h, kernel, variant, reps, int(flush_l2), C.byref(ms)), "time_kernel")
        return float(ms.value)

    TILE_INFO = ("max_local_vertices", "max_element_records", "max_entries", "max_items", "max_partial_buffers",
                 "smem_bytes_per_cta", "map_bytes", "threads_per_cta")

    def tile_info(self):
        """Statistics of the tile-assembly maps (valid after the first ``assemble(kernel=ASMK_TILE)``)."""
        info = np.zeros(8, dtype=np.int64)
        self._check(self._lib.glims_tile_info(self._h, N.as_lp(info)), "tile_info")
        return dict(zip(self.TILE_INFO, (int(v) for v in info)))

    def tile_config(self, threads_per_cta=0, chunk=0):
        self._check(self._lib.glims_tile_config(self._h, threads_per_cta, chunk), "tile_config")

    @property
    def nnzb(self):
        return int(self._lib.glims_nnzb(self._h))

    @property
    def nslots(self):
        return int(self._lib.glims_nslots(self._h))

    @property
    def launch_count(self):
        return int(self._lib.glims_launch_count(self._h))

    # -- multi-GPU ----------------------------------------------------------------------------
    @staticmethod
    def nccl_unique_id():
        buf = (C.c_char * 128)()
        rc = N.load().glims_nccl_unique_id(buf)
        if rc != 0:
            raise EngineError("ncclGetUniqueId failed")
        return bytes(buf)

    def comm_init(self, n_ranks, rank, unique_id):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self._lib.glims_comm_init(self._h, n_ranks, rank, buf), "comm_init")

    def set_p2p(self, on=True):
        """Halo exchange / allreduce transport: peer-memory windows over NVLink (True) or NCCL (False). Returns what is in use."""
        rc = self._lib.glims_set_p2p(self._h, int(bool(on)))
        if rc < 0:
            self._check(rc, "set_p2p")
        return bool(rc)

    def comm_bench(self, kind, reps=200):
        us = C.c_float()
        self._check(self._lib.glims_comm_bench(self._h, kind, reps, C.byref(us)), "comm_bench")
        return float(us.value)

    def set_halo(self, peers, send_ptr, send_idx, recv_ptr):
        peers = np.ascontiguousarray(peers, dtype=np.int32)
        send_ptr = np.ascontiguousarray(send_ptr, dtype=np.int64)
        send_idx = np.ascontiguousarray(send_idx, dtype=np.int32)
        recv_ptr = np.ascontiguousarray(recv_ptr, dtype=np.int64)
        self._check(self._lib.glims_set_halo(self._h, len(peers), N.as_ip(peers), N.as_lp(send_ptr),
                                             N.as_ip(send_idx), N.as_lp(recv_ptr)), "set_halo")
