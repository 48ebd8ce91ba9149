"""Test double for ``glimslib_b200.engine.Engine`` backed by the CPU oracle (test infrastructure: lives under tests/).

The product has no CPU path.  The container in which the reference tree exists has no GPU, and the GPU
box has no reference tree, so the *unmodified* reference scripts can only be exec'd against the drop-in API here with the
device engine swapped for this stand-in; the same API on the real engine is covered by tests/test_gpu_dropin.py."""
import numpy as np

from oracle import fem, solver as osolver


class OracleEngine:
    instances = []

    def __init__(self, coords, cells, cell_mat, device=0, n_owned=-1):
        self.coords, self.cells, self.cell_mat = np.asarray(coords), np.asarray(cells), np.asarray(cell_mat)
        self.dim = self.coords.shape[1]
        self.nb = self.dim + 1
        self.n_vertices, self.n_cells = len(self.coords), len(self.cells)
        self.n_owned = self.n_vertices
        self.ndof = self.n_vertices * self.nb
        self.table, self.dt, self.f_ext = None, 1.0, None
        self.bc_dofs, self.bc_vals = np.zeros(0, np.int64), np.zeros(0)
        self.x, self.x_prev = np.zeros(self.ndof), np.zeros(self.ndof)
        self.calls = []
        OracleEngine.instances.append(self)

    def _prob(self):
        t = self.table
        p = fem.Problem(self.coords, self.cells, self.cell_mat, fem.Materials(t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4]),
                        self.dt, bc_dofs=self.bc_dofs, bc_vals=self.bc_vals)
        p.f_ext = self.f_ext
        return p

    def set_materials(self, table): self.table = np.array(table, dtype=float); self.calls.append("set_materials")
    def set_dt(self, dt): self.dt = float(dt)
    def set_load(self, f): self.f_ext = None if f is None else np.array(f, dtype=float); self.calls.append("set_load")
    def set_dirichlet(self, dofs, vals):
        self.bc_dofs, self.bc_vals = np.asarray(dofs, np.int64).copy(), np.asarray(vals, float).copy()
        self.calls.append("set_dirichlet")
    def set_state(self, x): self.x = np.array(x, dtype=float).ravel()
    def set_prev(self, x): self.x_prev = np.array(x, dtype=float).ravel()
    def get_prev(self): return self.x_prev.copy()

    def get_state(self, out=None):
        if out is None:
            return self.x.copy()
        out[:] = self.x
        return out

    def step(self, n_steps=1, **opts):
        stats = []
        for _ in range(n_steps):
            st = {}
            x, its = osolver.newton(self._prob(), self.x, self.x_prev, rtol=opts.get("snes_rtol", 1e-9),
                                    atol=opts.get("snes_atol", 1e-10), linear="lu", stats=st)
            self.x, self.x_prev = x, x.copy()
            stats.append(dict(newton_its=its, krylov_its_c=0, krylov_its_u=0, krylov_its_mono=0, converged=1,
                              fnorm0=st["fnorm"][0], fnorm=st["fnorm"][-1], ms_total=0.0, ms_assembly=0.0, ms_krylov=0.0))
        self.last_stats = stats
        return stats

    def adjoint_gradient(self, n_steps, levels=(), level_targets=None, u_target=None, **opts):
        """Same contract as Engine.adjoint_gradient, from oracle/adjoint.py with one control per (material, coefficient)."""
        from oracle import adjoint
        prob = self._prob()
        n_mat = len(self.table)
        spec = [(name, m) for m in range(n_mat) for name in ("D", "rho", "gamma")]
        p = np.array([self.table[m][k] for m in range(n_mat) for k in (2, 3, 4)], dtype=float)
        d = self.coords.shape[1]
        targets = {"levels": {float(lv): np.asarray(level_targets)[i] for i, lv in enumerate(levels)}}
        if u_target is not None:
            targets["u"] = np.asarray(u_target, dtype=float).reshape(-1, d)
        J, g = adjoint.gradient(prob, self.x_prev.copy(), int(n_steps), targets, p, spec)
        xs = adjoint.forward(prob, self.x_prev.copy(), int(n_steps))
        self.x, self.x_prev = xs[-1].copy(), xs[-1].copy()
        self.calls.append("adjoint")
        return float(J), np.asarray(g).reshape(n_mat, 3)

    def cell_fields(self, vertex=False):
        raise NotImplementedError("derived fields run on the device only")

    def close(self):
        pass
