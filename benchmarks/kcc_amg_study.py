#!/usr/bin/env python
"""CPU study (scipy, no GPU): what an aggregation AMG would buy for the concentration block K_cc when diffusion is stiff
(config C2: dt*D/h^2 = 50, Jacobi-PCG needs ~316 iterations per step on the GPU).  Same ingredients as the library's K_uu
preconditioner, scalar version: greedy aggregation, piecewise-constant prolongator, Galerkin coarse operators,
Chebyshev(2)-Jacobi smoothing, V-cycle inside PCG.

    python benchmarks/kcc_amg_study.py [--n 300]
"""
import argparse
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))

from amg_partition_study import aggregate, pcg_iterations  # noqa: E402


def build(A):
    levels = []
    while True:
        A = A.tocsr()
        n = A.shape[0]
        dinv = 1.0 / A.diagonal()
        if n <= 400 or len(levels) >= 10:
            levels.append(dict(A=A, dense=np.linalg.inv(A.toarray())))
            return levels
        v = np.random.default_rng(0).standard_normal(n)
        for _ in range(12):
            v = dinv * (A @ v)
            lam = np.linalg.norm(v)
            v /= lam
        agg, na = aggregate(A.indptr, A.indices, n, np.zeros(n, bool), np.zeros(n, np.int64))
        P = sp.csr_matrix((np.ones(n), (np.arange(n), agg)), shape=(n, na))
        levels.append(dict(A=A, dinv=dinv, lmax=1.1 * lam, P=P))
        A = P.T @ A @ P


def smooth(L, b, x):
    lmax, lmin = L["lmax"], 0.1 * L["lmax"]
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma = theta / delta
    rho = 1.0 / sigma
    d = None
    for k in range(2):
        z = L["dinv"] * (b - L["A"] @ x if x is not None else b)
        if k == 0:
            d = z / theta
        else:
            rho_new = 1.0 / (2.0 * sigma - rho)
            d = rho_new * rho * d + 2.0 * rho_new / delta * z
            rho = rho_new
        x = d if x is None else x + d
    return x


def vcycle(levels, li, b):
    L = levels[li]
    if "dense" in L:
        return L["dense"] @ b
    x = smooth(L, b, None)
    x = x + L["P"] @ vcycle(levels, li + 1, L["P"].T @ (b - L["A"] @ x))
    return smooth(L, b, x)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=300)
    args = ap.parse_args()
    from glimslib_b200 import workloads as W
    from oracle import fem
    w = W.c2_2d_1m(args.n)
    w["table"][:, 2] = 1e-4 * (707.0 / args.n) ** 2      # keep dt*D/h^2 of the full-size C2 mesh (= 50)
    t = w["table"]
    prob = fem.Problem(w["mesh"].coords, w["mesh"].cells, w["cell_mat"],
                       fem.Materials(t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4]), w["dt"])
    x = w["x0"].copy()
    F, J = fem.assemble(prob, x, x)
    ic = (np.arange(prob.ndof) % 3) == 2
    Kcc = J[ic][:, ic].tocsr()
    b = -F[ic]
    dinv = 1.0 / Kcc.diagonal()
    levels = build(Kcc)
    out = {"n": args.n, "n_vertices": int(Kcc.shape[0]), "dt_D_over_h2": float(t[0, 2] * args.n ** 2),
           "levels": [int(L["A"].shape[0]) for L in levels],
           "pcg_iterations": {"jacobi": pcg_iterations(Kcc, b, lambda r: dinv * r, maxit=2000),
                              "amg_vcycle": pcg_iterations(Kcc, b, lambda r: vcycle(levels, 0, r), maxit=2000)}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
