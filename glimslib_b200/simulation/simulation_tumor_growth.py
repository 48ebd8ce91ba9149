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

    def init_postprocess(self, output_dir=config.output_dir_simulation_tmp):
        self.postprocess = PostProcessTumorGrowth(self.results, self.params, output_dir=output_dir,
                                                  engine=getattr(getattr(self, "solver", None), "_engine", None))
        self.postprocess._form = getattr(getattr(getattr(self, "solver", None), "problem", None), "form", None)
