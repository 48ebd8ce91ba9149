"""Base class of the drop-in simulations: same public API and time-loop semantics as
``glimslib/simulation/simulation_base.py`` (``FenicsSimulation``), driving the B200 backend.

Public surface kept: ``__init__(mesh, time_dependent)``, ``setup_global_parameters``,
``setup_model_parameters``, ``run(keep_nth, save_method, clear_all, plot, output_dir)`` returning
``self.solution``, ``reload_from_hdf5``; attributes ``mesh, functionspace, subdomains, bcs, params, results,
solution, solver``.  Time loop semantics (simulation_base.py:253-317): t=0 record of the initial value, loop
while ``t <= sim_time - 1e-5``, a failing solve logs a warning and stops the loop (no exception escapes),
record every ``keep_nth`` steps, ``u_previous.assign(solution)``, then ``solution_timeseries.h5``.
"""
import logging
import os

import numpy as np
from abc import ABC, abstractmethod

from glimslib_b200 import fenics_local as fenics
from glimslib_b200.simulation import config
from glimslib_b200.simulation_helpers.helper_classes import (BoundaryConditions, FunctionSpace, Parameters, Plotting,
                                                             Results, SubDomains)


class FenicsSimulation(ABC):
    def __init__(self, mesh, time_dependent=True):
        self.logger = logging.getLogger(__name__)
        self.mesh = mesh
        self.geometric_dimension = mesh.geometry().dim
        self.time_dependent = time_dependent
        self.projection_parameters = {"solver_type": "cg", "preconditioner_type": "amg"}
        self.functionspace = FunctionSpace(mesh, projection_parameters=self.projection_parameters)
        self._engine_cache = {}
        self._define_model_params()

    # -- to be provided by the model classes ------------------------------------------------------
    @abstractmethod
    def _define_model_params(self):
        self.required_params, self.optional_params = [], []

    @abstractmethod
    def _setup_functionspace(self):
        ...

    @abstractmethod
    def _setup_problem(self, u_previous):
        """Must leave ``self.solution`` and ``self.solver`` (object with ``solve()``) behind."""

    @abstractmethod
    def run_for_adjoint(self, parameters):
        ...

    # -- setup --------------------------------------------------------------------------------------
    def setup_global_parameters(self, label_function=None, subdomains=None, domain_names=None, boundaries=None,
                                dirichlet_bcs=None, von_neumann_bcs=None):
        self.logger.info("-- Setting up global parameters")
        self.geometric_dimension = self.mesh.geometry().dim
        self.subdomains = SubDomains(self.mesh)
        self.subdomains.setup_subdomains(label_function=label_function, subdomains=subdomains, replace=False)
        self.subdomains.setup_boundaries(tissue_map=domain_names, boundary_fct_dict=boundaries)
        self.subdomains.setup_measures()
        self._setup_functionspace()
        self.bcs = BoundaryConditions(self.functionspace, self.subdomains)
        self.bcs.setup_dirichlet_boundary_conditions(dirichlet_bcs)
        self.bcs.setup_von_neumann_boundary_conditions(von_neumann_bcs)

    def setup_model_parameters(self, iv_expression, **kwargs):
        self._define_model_params()
        self.params = Parameters(self.functionspace, self.subdomains, time_dependent=self.time_dependent)
        self.params.set_initial_value_expressions(iv_expression)
        self.params.define_required_params(self.required_params)
        self.params.define_optional_params(self.optional_params)
        self.params.init_parameters(kwargs)

    def _update_expressions(self, time):
        self.params.time_update_parameters(time)
        self.bcs.time_update_bcs(time, kind="dirichlet")
        self.bcs.time_update_bcs(time, kind="von-neumann")

    # -- the timed path -------------------------------------------------------------------------------
    def _record(self, t, time_step, recording_step, field, save_method, plot, function=None):
        self.results.add_to_results(t, time_step, recording_step, field)
        self.results.save_solution(recording_step, t, function=function, method=save_method)
        if plot:
            self.plotting.plot_all(recording_step)

    def run(self, keep_nth=1, save_method="xdmf", clear_all=False, plot=True,
            output_dir=config.output_dir_simulation_tmp):
        if self.geometric_dimension == 3:
            plot = False
        self.logger.info("-- Computing solutions: ")
        self.results = Results(self.functionspace, self.subdomains, output_dir=output_dir)
        self.results.save_solution_start(method=save_method, clear_all=clear_all)
        self.plotting = Plotting(self.results, output_dir=os.path.join(output_dir, "plots"))
        u_previous = self.params.create_initial_value_function()
        self._setup_problem(u_previous)

        if not self.time_dependent:
            self.solver.solve()
            self._record(0, 0, 0, self.solution, save_method, plot)
            u_previous.vector()[:] = self.solution.vector()
        else:
            t, step, rec = 0.0, 0, 0
            self._update_expressions(t)
            self._record(0, 0, rec, u_previous, save_method, plot, function=u_previous)
            dt = float(self.params.sim_time_step)
            while t <= self.params.sim_time - 1e-5:
                t += dt
                step += 1
                self._update_expressions(t)
                self.logger.info("    - solving for time = %.2f / %.2f" % (t, self.params.sim_time))
                try:
                    self.solver.solve()
                except fenics.SolverNotConverged as e:
                    # simulation_base.py:301-305: a failed solve is swallowed -- warn, stop, still save.
                    # Anything else (missing library, no GPU, CUDA fault) is NOT a convergence failure and
                    # propagates: the product path has no fallback and must fail loudly.
                    self.logger.warning("    - Solver did not converge -- will shutdown simulation (%s)" % e)
                    break
                if step % keep_nth == 0:
                    rec += 1
                    self._record(t, step, rec, self.solution, save_method, plot)
                u_previous.assign(self.solution)
                if hasattr(self.solver, "note_previous_assigned"):
                    self.solver.note_previous_assigned(u_previous, self.solution)

        self.results.save_solution_end(method=save_method)
        self.results.save_solution_hdf5()
        return self.solution

    def _update_mesh_displacements(self, displacement):
        """Adds a nodal displacement to the mesh coordinates (``fenics.ALE.move``, simulation_base.py:228-234; repeated calls
        accumulate).  Host mesh only -- output in the deformed configuration; the device engine keeps the configuration it
        was built on, exactly as the reference's solver does between calls."""
        u = displacement.node_values() if hasattr(displacement, "node_values") else displacement.values()
        self.mesh.coords += np.asarray(u).reshape(self.mesh.coords.shape)

    def reload_from_hdf5(self, path_to_hdf5, output_dir=config.output_dir_simulation_tmp):
        self.logger.info("-- Reloading from hdf5: ")
        self.results = Results(self.functionspace, self.subdomains, output_dir=output_dir)
        self.results.data.load_from_hdf5(path_to_hdf5)
        self.plotting = Plotting(self.results, output_dir=os.path.join(output_dir, "plots"))
