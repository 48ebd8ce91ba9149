"""``TumorGrowth``: mechanically-coupled reaction-diffusion tumour growth (drop-in for
``glimslib/simulation/simulation_tumor_growth.py``).

The weak form of stg:110-120 is not re-declared in UFL: ``_setup_problem`` turns the parameters into the
per-material table ``(mu, lambda, D, rho, gamma)`` and the load terms that the CUDA element kernels evaluate
in closed form (exact for P1 with per-cell-constant coefficients).
"""
import numpy as np

from glimslib_b200 import fenics_local as fenics
from glimslib_b200.simulation import config
from glimslib_b200.simulation.simulation_base import FenicsSimulation
from glimslib_b200.simulation_helpers import math_linear_elasticity as mle
from glimslib_b200.simulation_helpers.helper_classes import DiscontinuousScalar, PostProcessTumorGrowth


def _per_label(param, labels):
    """Scalar or DiscontinuousScalar -> value for each label id in ``labels``."""
    if isinstance(param, DiscontinuousScalar):
        return np.array([param.value_for_label(l) for l in labels], dtype=np.float64)
    if isinstance(param, fenics.Constant):
        return np.full(len(labels), float(param))
    return np.full(len(labels), float(param))


class TumorGrowth(FenicsSimulation):
    def __init__(self, mesh, time_dependent=True):
        super().__init__(mesh, time_dependent=time_dependent)
        self.units = {"motility": "m^2/s", "Emodulus": "N/m^2", "none": "", "growth_rate": "1/s"}

    def _setup_functionspace(self):
        """Mixed [P1^d, P1]: sub-space 0 displacement, 1 concentration (stg:67-72)."""
        cell = self.mesh.ufl_cell()
        element = fenics.MixedElement([fenics.VectorElement("Lagrange", cell, 1), fenics.FiniteElement("Lagrange", cell, 1)])
        self.functionspace.init_function_space(element, {0: "displacement", 1: "concentration"})

    def _define_model_params(self):
        self.required_params = ["diffusion", "coupling", "proliferation", "E", "poisson"]
        self.optional_params = []

    # -- material table: one row per label id present in the mesh ------------------------------------
    def _material_rows(self, labels):
        p = self.params
        E, nu = _per_label(p.E, labels), _per_label(p.poisson, labels)
        return np.stack([mle.compute_mu(E, nu), mle.compute_lambda(E, nu), _per_label(p.diffusion, labels),
                         _per_label(p.proliferation, labels), _per_label(p.coupling, labels)], axis=1)

    def _setup_problem(self, u_previous):
        dim = self.geometric_dimension
        lab = np.asarray(self.subdomains.subdomains.array())
        labels, cell_mat = np.unique(lab, return_inverse=True)
        table = self._material_rows(labels)
        if not hasattr(self, "body_force"):
            self.body_force = fenics.Constant(np.zeros(dim))          # stg:91-92
        if not hasattr(self, "source_term"):
            self.source_term = fenics.Constant(0.0)                    # stg:95-96
        W = self.functionspace.function_space
        self.solution = fenics.Function(W, name="solution_function")   # fresh zero Function: first Newton guess (stg:100)
        self.solution.label = "solution_function"
        self.logger.info("    - Using non-linear solver")
        neumann = self.bcs.neumann_terms(0) + self.bcs.neumann_terms(1)
        F = fenics.CoupledRDMechanicsForm(W, self.solution, u_previous, cell_mat.reshape(-1), table,
                                          dt=float(self.params.sim_time_step), body_force=self.body_force,
                                          source=self.source_term, neumann=neumann, engine_cache=self._engine_cache,
                                          table_fn=lambda: self._material_rows(labels))
        problem = fenics.NonlinearVariationalProblem(F, self.solution, bcs=getattr(self.bcs, "dirichlet_bcs", []), J=None)
        solver = fenics.NonlinearVariationalSolver(problem)
        prm = solver.parameters
        prm["nonlinear_solver"] = "snes"
        prm["snes_solver"]["report"] = False
        self.solver = solver

    def run_for_adjoint(self, parameters, output_dir=config.output_dir_simulation_tmp):
        """Forward run with updated (diffusion, proliferation, coupling) -- stg:142-156 (no tape here)."""
        self.params.diffusion, self.params.proliferation, self.params.coupling = parameters
        self.run(keep_nth=1, save_method=None, clear_all=False, plot=False, output_dir=output_dir)
        return self.solution

    def run_for_adjoint2(self, parameters, output_dir=config.output_dir_simulation_tmp):
        self.params.diffusion, self.params.proliferation = parameters
        self.run(keep_nth=1, save_method=None, clear_all=False, plot=False, output_dir=output_dir)
        return self.solution

    # -- inverse problem: value and gradient of the image misfit -------------------------------------
    def _n_time_steps(self):
        t, dt, n = 0.0, float(self.params.sim_time_step), 0
        while t <= self.params.sim_time - 1e-5:          # the loop of `run` (simulation_base.py:285-312)
            t += dt
            n += 1
        return n

    def _set_controls(self, parameters):
        self.params.diffusion, self.params.proliferation, self.params.coupling = parameters

    def _controls_from_material_gradient(self, grad, labels):
        """(dJ/d diffusion, dJ/d proliferation, dJ/d coupling): the three controls are uniform over the tissues."""
        return np.asarray(grad).sum(axis=0)

    def misfit_gradient(self, parameters, threshold_levels, threshold_targets, displacement_target=None):
        """What the reference obtains from dolfin-adjoint for ``run_for_adjoint`` (image_based_optimization.py:660-767): the
        misfit J = sum_l |th_l(c_N) - target_l|^2 + |u_N - u_target|^2 (mass-weighted L2, th_l the smooth threshold at level l,
        :1404-1407) of the run with the given controls, and dJ/d(controls) -- one forward run and one backward sweep on the
        device (``glims_adjoint``), time-independent coefficients.  Targets are scalar P1 functions (or vertex arrays), the
        displacement target a vector P1 function (or [n_vertices, dim] array).  Returns (J, gradient)."""
        self._set_controls(parameters)
        u_previous = self.params.create_initial_value_function()
        self._update_expressions(0.0)
        self._setup_problem(u_previous)
        arr = lambda f: np.asarray(f.vector().get_local() if hasattr(f, "vector") else f, dtype=np.float64)
        nv = self.mesh.num_vertices()
        targets = np.stack([arr(f).reshape(nv) for f in threshold_targets]) if len(threshold_levels) else np.zeros((0, nv))
        ut = None if displacement_target is None else arr(displacement_target).reshape(nv, self.geometric_dimension)
        J, grad = self.solver.adjoint_gradient(self._n_time_steps(), threshold_levels, targets, ut)
        labels = np.unique(np.asarray(self.subdomains.subdomains.array()))
        return J, self._controls_from_material_gradient(grad, labels)

    def init_postprocess(self, output_dir=config.output_dir_simulation_tmp):
        self.postprocess = PostProcessTumorGrowth(self.results, self.params, output_dir=output_dir,
                                                  engine=getattr(getattr(self, "solver", None), "_engine", None))
        self.postprocess._form = getattr(getattr(getattr(self, "solver", None), "problem", None), "form", None)
