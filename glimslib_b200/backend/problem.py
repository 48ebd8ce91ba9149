"""``NonlinearVariationalProblem`` / ``NonlinearVariationalSolver`` of the drop-in seam.

The reference builds a UFL residual ``F`` and its derivative ``J`` (simulation_tumor_growth.py:110-124) and hands
them to DOLFIN (:126-130); ``solver.solve()`` is then called once per time step (simulation_base.py:302).  Here
``F`` is a :class:`CoupledRDMechanicsForm` -- a *description* of that same residual (mesh, per-cell materials,
dt, loads, u_previous) -- and ``solve()`` drives the CUDA engine: element kernels, fixed-pattern assembly,
Dirichlet elimination, Krylov solves and the Newton loop all run on the device (include/glims_b200.h).
"""
import logging

import numpy as np

from ..engine import Engine, SolverNotConverged  # noqa: F401  (no CPU fallback: import fails without the library)
from . import core

log = logging.getLogger(__name__)


class CoupledRDMechanicsForm:
    """Residual of stg:110-120 for per-cell-constant coefficients.

    table[m] = (mu, lambda, D, rho, gamma) for compact material index m; cell_mat[c] = m.
    neumann: list of (facet ids, subspace_id, value) on exterior facets (helper_classes.py:861-908).
    """

    def __init__(self, function_space, solution, u_previous, cell_mat, table, dt, body_force=None,
                 source=None, neumann=(), engine_cache=None, table_fn=None):
        self.V, self.solution, self.u_previous = function_space, solution, u_previous
        self.cell_mat = np.ascontiguousarray(cell_mat, dtype=np.int32)
        self.table = np.ascontiguousarray(table, dtype=np.float64)
        self.dt, self.body_force, self.source, self.neumann = float(dt), body_force, source, list(neumann)
        self.engine_cache = engine_cache if engine_cache is not None else {}
        self.table_fn = table_fn          # re-evaluates the per-label coefficients (time-dependent parameters)

    @staticmethod
    def _coef_sig(obj):
        """What a coefficient object currently evaluates to, cheaply: its time stamp `.t` (set by
        simulation_base._update_expressions every step), its user parameters and, for Constants, its values."""
        if obj is None:
            return None
        sig = [id(obj)]
        if isinstance(obj, core.Constant):
            sig.append(obj.values().tobytes())          # a `.t` stamped on a Constant changes nothing
        elif isinstance(obj, core.Expression):
            sig.append(repr(sorted((k, repr(v)) for k, v in obj.__dict__.get("_params", {}).items())))
            sig.append(repr(obj.__dict__.get("t")))     # Python subclasses whose eval() reads self.t
        elif isinstance(obj, (int, float)):
            sig.append(float(obj))
        else:
            try:
                sig.append(("t", float(obj.t)))
            except Exception:
                pass
        return tuple(sig)

    def coefficient_signature(self):
        """Changes whenever a load term (body force, source, von-Neumann value) may evaluate differently than at the last
        solve -- the reference's UFL form re-evaluates them at every assembly (stg:110-120)."""
        return (self._coef_sig(self.body_force), self._coef_sig(self.source),
                tuple(self._coef_sig(v) for _, _, v in self.neumann))

    def load_vector(self):
        """f_ext so that F = F_int - f_ext: body force (stg:112), RD source (stg:119), Neumann terms (stg:113,120)."""
        mesh = self.V.mesh()
        d = mesh.dim
        nb = d + 1
        have = False
        f = np.zeros((mesh.num_vertices(), nb))
        X = mesh.coords[mesh.cells]
        vol = np.abs(np.linalg.det(X[:, 1:] - X[:, :1])) / {2: 2.0, 3: 6.0}[d]
        bf = None if self.body_force is None else np.asarray(self.body_force.values() if hasattr(self.body_force, "values") else self.body_force, float)
        if bf is not None and np.any(bf != 0):
            have = True
            for a in range(nb):
                np.add.at(f[:, :d], mesh.cells[:, a], (vol / nb)[:, None] * bf[None, :])
        s = None if self.source is None else float(self.source)
        if s:
            have = True
            for a in range(nb):
                np.add.at(f[:, d], mesh.cells[:, a], self.dt * s * vol / nb)
        if self.neumann:
            fv, _, fc = mesh.facets()
            for fids, subspace, value in self.neumann:
                if len(fids) == 0:
                    continue
                verts = fv[fids]
                Xf = mesh.coords[verts]
                if d == 2:
                    area = np.linalg.norm(Xf[:, 1] - Xf[:, 0], axis=1)
                else:
                    area = 0.5 * np.linalg.norm(np.cross(Xf[:, 1] - Xf[:, 0], Xf[:, 2] - Xf[:, 0]), axis=1)
                mid = Xf.mean(axis=1)
                g = value.eval_points(mid) if hasattr(value, "eval_points") else np.broadcast_to(np.atleast_1d(float(value)), (len(fids), 1))
                if np.any(g != 0):
                    have = True
                w = area / d
                if subspace == 0:
                    for k in range(d):
                        np.add.at(f[:, :d], verts[:, k], w[:, None] * g)
                else:
                    Dc = self.table[self.cell_mat[fc[fids, 0]], 2]
                    for k in range(d):
                        np.add.at(f[:, d], verts[:, k], self.dt * Dc * g[:, 0] * w)
        return f.ravel() if have else None


class _Params(dict):
    """dict with attribute access and nested defaults, like DOLFIN's Parameters."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


class NonlinearVariationalProblem:
    def __init__(self, F, u, bcs=None, J=None, **kw):
        if not isinstance(F, CoupledRDMechanicsForm):
            raise NotImplementedError("the B200 backend solves the coupled RD-mechanics form only "
                                      "(simulation_tumor_growth.py:110-124); got %r" % type(F))
        self.form, self.u = F, u
        self.bcs = list(bcs) if bcs is not None else []


class NonlinearVariationalSolver:
    def __init__(self, problem):
        self.problem = problem
        snes = _Params(report=True, relative_tolerance=1e-9, absolute_tolerance=1e-10, solution_tolerance=1e-16,
                       maximum_iterations=50, error_on_nonconvergence=True, linear_solver="default",
                       preconditioner="default", line_search="basic", method="default",
                       krylov_solver=_Params(relative_tolerance=1e-10, absolute_tolerance=1e-300,
                                             maximum_iterations=20000))
        newton = _Params(report=True, relative_tolerance=1e-9, absolute_tolerance=1e-10, maximum_iterations=50,
                         error_on_nonconvergence=True, linear_solver="default", preconditioner="default",
                         krylov_solver=_Params(relative_tolerance=1e-10, absolute_tolerance=1e-300,
                                               maximum_iterations=20000))
        self.parameters = _Params(nonlinear_solver="newton", snes_solver=snes, newton_solver=newton,
                                  # B200 backend extensions
                                  b200=_Params(solver="block_tri", preconditioner="amg", assembly="rows",
                                               lag_mechanics=True, device=0))
        self._engine = None
        self._pushed_version = (None, None)
        self.last_stats = None

    # -------------------------------------------------------------------------------------------
    def _get_engine(self):
        form = self.problem.form
        mesh = form.V.mesh()
        key = (id(mesh), mesh.num_cells(), hash(form.cell_mat.tobytes()), int(self.parameters.b200.device))
        eng = form.engine_cache.get(key)
        if eng is None:
            for e in form.engine_cache.values():
                e.close()
            form.engine_cache.clear()
            eng = Engine(mesh.coords, mesh.cells, form.cell_mat, device=int(self.parameters.b200.device))
            form.engine_cache[key] = eng
        return eng

    def _configure(self, eng):
        form = self.problem.form
        eng.set_materials(form.table)
        eng.set_dt(form.dt)
        eng.set_load(form.load_vector())
        self._coef_sig = form.coefficient_signature()
        self._table_pushed = form.table.copy()
        self._configured = True

    def _refresh_coefficients(self, eng):
        """Time-dependent coefficients: simulation_base._update_expressions(t) stamps `.t` on parameters and boundary
        values before every solve and the reference's form picks the new values up at assembly time.  Here the material
        table and the pre-integrated load vector live on the device, so they are re-pushed when (and only when) something
        they were built from changed."""
        form = self.problem.form
        if form.table_fn is not None:
            table = np.ascontiguousarray(form.table_fn(), dtype=np.float64)
            if table.shape != self._table_pushed.shape or not np.array_equal(table, self._table_pushed):
                form.table = table
                eng.set_materials(table)
                self._table_pushed = table.copy()
        sig = form.coefficient_signature()
        if sig != self._coef_sig:
            eng.set_load(form.load_vector())
            self._coef_sig = sig

    def _bc_arrays(self):
        dofs, vals = [], []
        for bc in self.problem.bcs:
            bc.update()
            dofs.append(bc.dofs)
            vals.append(bc.values)
        if not dofs:
            return np.zeros(0, dtype=np.int64), np.zeros(0)
        dofs, vals = np.concatenate(dofs), np.concatenate(vals)
        # later conditions win on shared dofs, as successive DirichletBC.apply calls would
        _, last = np.unique(dofs[::-1], return_index=True)
        keep = len(dofs) - 1 - last
        return dofs[keep], vals[keep]

    def _prepare(self):
        """Engine (created and configured on first use), current coefficients, Dirichlet data, solver options, and the
        host -> device copies of whatever changed on the host since the last call.  Returns (engine, options, nl-params)."""
        from .. import _native as N
        form, prm = self.problem.form, self.parameters
        if self._engine is None:
            self._engine = self._get_engine()
            self._configure(self._engine)
            self._bc_sig = None
        eng = self._engine
        self._refresh_coefficients(eng)
        dofs, vals = self._bc_arrays()
        sig = (hash(dofs.tobytes()), hash(vals.tobytes()))
        if sig != self._bc_sig:
            eng.set_dirichlet(dofs, vals)
            self._bc_sig = sig
        nl = prm[prm.nonlinear_solver + "_solver"] if prm.nonlinear_solver in ("snes", "newton") else prm.snes_solver
        ks = nl.krylov_solver
        tight = nl.linear_solver in ("lu", "mumps", "superlu", "umfpack", "petsc")
        opts = dict(snes_rtol=float(nl.relative_tolerance), snes_atol=float(nl.absolute_tolerance),
                    max_newton=int(nl.maximum_iterations),
                    ksp_rtol=1e-13 if tight else float(ks.relative_tolerance), ksp_atol=float(ks.absolute_tolerance),
                    max_krylov=int(ks.maximum_iterations),
                    solver=N.SOLVER_MONO_GMRES if prm.b200.solver == "mono_gmres" else N.SOLVER_BLOCK_TRI,
                    pc=N.PC_AMG if prm.b200.preconditioner == "amg" else N.PC_JACOBI,
                    asm_kernel={"gather": N.ASMK_GATHER, "atomic": N.ASMK_ATOMIC, "tile": N.ASMK_TILE}.get(prm.b200.assembly, N.ASMK_ROWS),
                    lag_mechanics=int(bool(prm.b200.lag_mechanics)))
        u, up = self.problem.u, form.u_previous
        # host -> device only for what changed on the host since the last solve
        if self._pushed_version[1] != (id(up), up.version):
            eng.set_prev(up._x)
        if self._pushed_version[0] != (id(u), u.version):
            eng.set_state(u._x)
        return eng, opts, nl

    def adjoint_gradient(self, n_steps, levels, level_targets, u_target=None):
        """``n_steps`` forward steps from (u_previous, u) and the discrete adjoint of the reference's image misfit of the final
        state (image_based_optimization.py:660-700) on the device (``glims_adjoint``).  Returns ``(J, grad)`` with
        ``grad[m] = (dJ/dD, dJ/drho, dJ/dgamma)`` of material row ``m`` of the form's table; ``u`` holds the final state."""
        eng, opts, _ = self._prepare()
        J, grad = eng.adjoint_gradient(int(n_steps), list(levels), np.asarray(level_targets, dtype=np.float64),
                                       None if u_target is None else np.asarray(u_target, dtype=np.float64), **opts)
        u = self.problem.u
        eng.get_state(out=u._x)
        u._touch()
        self._pushed_version = ((id(u), u.version), None)
        self._device_prev_is = u.version
        return J, grad

    def solve(self):
        """One Newton-Krylov solve on the device; returns (iterations, converged) like DOLFIN."""
        eng, opts, nl = self._prepare()
        u = self.problem.u
        try:
            stats = eng.step(1, **opts)[0]
        except SolverNotConverged:
            self.last_stats = eng.last_stats[0] if getattr(eng, "last_stats", None) else None
            if nl.error_on_nonconvergence:
                raise
            stats = self.last_stats
        self.last_stats = stats
        eng.get_state(out=u._x)
        u._touch()
        # the device did u_previous.assign(solution) already (glims_step); remember what it holds
        self._pushed_version = ((id(u), u.version), None)
        self._device_prev_is = u.version
        if nl.report:
            log.info("    Newton its %d, Krylov its c/u %d/%d, |F| %.3e -> %.3e", stats["newton_its"],
                     stats["krylov_its_c"], stats["krylov_its_u"], stats["fnorm0"], stats["fnorm"])
        return stats["newton_its"], bool(stats["converged"])

    def note_previous_assigned(self, u_previous, solution):
        """Called by the time loop after ``u_previous.assign(solution)``: the device already holds that copy,
        so the next solve need not upload it again."""
        if getattr(self, "_device_prev_is", None) == solution.version:
            self._pushed_version = (self._pushed_version[0], (id(u_previous), u_previous.version))
