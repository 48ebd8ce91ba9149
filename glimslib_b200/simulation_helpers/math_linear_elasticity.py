"""Lame parameters (``glimslib/simulation_helpers/math_linear_elasticity.py:6-10``). The strain / stress /
growth-strain one-liners of :12-17,32-33 are evaluated inside the CUDA element kernels
(``csrc/kernels.cu: k_assemble_*``); numpy versions for per-cell post-processing live here."""
import numpy as np


def compute_mu(young_modulus, poisson_ratio):
    return young_modulus / (2.0 * (1.0 + poisson_ratio))


def compute_lambda(young_modulus, poisson_ratio):
    return young_modulus * poisson_ratio / ((1.0 + poisson_ratio) * (1.0 - 2.0 * poisson_ratio))


def compute_strain(grad_u):
    """sym(grad u) for an array of per-cell displacement gradients [..., d, d]."""
    return 0.5 * (grad_u + np.swapaxes(grad_u, -1, -2))


def compute_stress(grad_u, mu, lmbda):
    eps = compute_strain(grad_u)
    d = eps.shape[-1]
    tr = np.trace(eps, axis1=-2, axis2=-1)
    return 2.0 * np.asarray(mu)[..., None, None] * eps + (np.asarray(lmbda) * tr)[..., None, None] * np.eye(d)


def compute_growth_induced_strain(conc, coupling_constant, dim):
    return (np.asarray(conc) * coupling_constant)[..., None, None] * np.eye(dim)
