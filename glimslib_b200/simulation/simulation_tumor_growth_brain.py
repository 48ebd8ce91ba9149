"""``TumorGrowthBrain``: the per-tissue variant (drop-in for
``glimslib/simulation/simulation_tumor_growth_brain.py``): scalar parameters per tissue name
(GM, WM, CSF, Ventricles), D = rho = 0 in CSF / Ventricles / outside (:94-102), one scalar ``coupling``
everywhere, hard-coded 'outside' material E=10e3, nu=0.45 (:37-38), and no von-Neumann terms (:91,105).
It produces the same element kernel inputs as ``TumorGrowth`` with a per-tissue table.
"""
import numpy as np

from glimslib_b200.simulation import config
from glimslib_b200.simulation.simulation_tumor_growth import TumorGrowth
from glimslib_b200.simulation_helpers import math_linear_elasticity as mle
from glimslib_b200.simulation_helpers.helper_classes import PostProcessTumorGrowthBrain


class TumorGrowthBrain(TumorGrowth):
    def _define_model_params(self):
        self.required_params = ["E_GM", "E_WM", "E_CSF", "E_VENT", "nu_GM", "nu_WM", "nu_CSF", "nu_VENT",
                                "D_GM", "D_WM", "rho_GM", "rho_WM", "coupling"]
        self.optional_params = []

    def _material_rows(self, labels):
        p, sd = self.params, self.subdomains
        g = float(p.coupling)
        by_name = {
            "GM": (float(p.E_GM), float(p.nu_GM), float(p.D_GM), float(p.rho_GM), g),
            "WM": (float(p.E_WM), float(p.nu_WM), float(p.D_WM), float(p.rho_WM), g),
            "CSF": (float(p.E_CSF), float(p.nu_CSF), 0.0, 0.0, g),
            "Ventricles": (float(p.E_VENT), float(p.nu_VENT), 0.0, 0.0, g),
            "outside": (10e3, 0.45, 0.0, 0.0, g),
        }
        rows = []
        for lab in labels:
            name = sd.tissue_id_name_map.get(int(lab))
            if name not in by_name:
                # the reference integrates only over dx(<the five names>): other labels contribute nothing but
                # the mass term (:93,98); a zero-stiffness region would be singular, so refuse loudly
                raise ValueError("TumorGrowthBrain: label %r (%r) is not one of GM/WM/CSF/Ventricles/outside" % (lab, name))
            E, nu, D, rho, gam = by_name[name]
            rows.append((mle.compute_mu(E, nu), mle.compute_lambda(E, nu), D, rho, gam))
        return np.asarray(rows, dtype=np.float64)

    def _setup_problem(self, u_previous):
        if not hasattr(self, "rd_source_term"):
            from glimslib_b200 import fenics_local as fenics
            self.rd_source_term = fenics.Constant(0)
        self.source_term = self.rd_source_term
        # no von-Neumann terms in this variant
        saved = getattr(self.bcs, "von_neumann_bcs", None)
        self.bcs.von_neumann_bcs = {}
        try:
            super()._setup_problem(u_previous)
        finally:
            if saved is not None:
                self.bcs.von_neumann_bcs = saved
            else:
                del self.bcs.von_neumann_bcs

    def run_for_adjoint(self, parameters, output_dir=config.output_dir_simulation_tmp):
        """Forward run with updated (D_WM, D_GM, rho_WM, rho_GM, coupling) -- brain:127-145 (no tape here)."""
        p = self.params
        p.D_WM, p.D_GM, p.rho_WM, p.rho_GM, p.coupling = parameters
        self.run(keep_nth=1, save_method=None, clear_all=False, plot=False, output_dir=output_dir)
        return self.solution

    def _set_controls(self, parameters):
        p = self.params
        p.D_WM, p.D_GM, p.rho_WM, p.rho_GM, p.coupling = parameters

    def _controls_from_material_gradient(self, grad, labels):
        """(dJ/dD_WM, dJ/dD_GM, dJ/drho_WM, dJ/drho_GM, dJ/dcoupling), the order of run_for_adjoint (brain:127-145)."""
        grad = np.asarray(grad)
        row = {self.subdomains.tissue_id_name_map.get(int(lab)): k for k, lab in enumerate(labels)}
        pick = lambda name, col: float(grad[row[name], col]) if name in row else 0.0
        return np.array([pick("WM", 0), pick("GM", 0), pick("WM", 1), pick("GM", 1), float(grad[:, 2].sum())])

    def init_postprocess(self, output_dir=config.output_dir_simulation_tmp):
        self.postprocess = PostProcessTumorGrowthBrain(self.results, self.params, output_dir=output_dir,
                                                       engine=getattr(getattr(self, "solver", None), "_engine", None))
        self.postprocess._form = getattr(getattr(getattr(self, "solver", None), "problem", None), "form", None)
