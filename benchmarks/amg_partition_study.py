#!/usr/bin/env python
"""CPU study (numpy/scipy, no GPU): how much of the multi-GPU iteration penalty of the K_uu solve comes from the
rank-local AMG hierarchy, and what recovers it.

The library's preconditioner (csrc/amg.cu) is an aggregation AMG with rigid-body-mode prolongators, Chebyshev(2)
block-Jacobi smoothing and Galerkin coarse operators.  On a partitioned mesh every rank aggregates its own vertices and
its coarse operators drop the couplings to other ranks' aggregates (ghost columns are not in the coarse space); only the
fine-level smoother is global.  Measured on B200s (profiles/): 15.7 PCG iterations per step on one GPU, 22.7 on eight
(10M tets), and +50 % at 50M tets.

This script rebuilds that preconditioner in scipy on the C4 operator of a small voxel ellipsoid and counts PCG
iterations for one right-hand side with
  global      one hierarchy over the whole mesh (= one GPU)
  local-drop  aggregates confined to RCB parts, cross-part couplings dropped on every coarse level (= the library now)
  local-keep  the same aggregates, but the true Galerkin operators (cross-part couplings kept: needs a ghost exchange on
              the coarse levels, no change to the aggregation)
  local-keep1 couplings kept on level 1 only

    python benchmarks/amg_partition_study.py [--grid 40] [--parts 8]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def aggregate(indptr, indices, n, excluded, part):
    """Greedy Vanek-style aggregation (csrc/amg.cu: aggregate), restricted to neighbours of the same part."""
    agg = -np.ones(n, dtype=np.int64)
    na = 0
    ok = lambda i, j: j != i and not excluded[j] and part[j] == part[i]
    for i in range(n):
        if excluded[i] or agg[i] >= 0:
            continue
        nb = [j for j in indices[indptr[i]:indptr[i + 1]] if ok(i, j)]
        if not nb or any(agg[j] >= 0 for j in nb):
            continue
        agg[i] = na
        agg[nb] = na
        na += 1
    agg1 = agg.copy()
    for i in range(n):
        if excluded[i] or agg1[i] >= 0:
            continue
        for j in indices[indptr[i]:indptr[i + 1]]:
            if ok(i, j) and agg1[j] >= 0:
                agg[i] = agg1[j]
                break
    for i in range(n):
        if excluded[i] or agg[i] >= 0:
            continue
        agg[i] = na
        for j in indices[indptr[i]:indptr[i + 1]]:
            if ok(i, j) and agg[j] < 0:
                agg[j] = na
        na += 1
    return agg, na


def prolongator(X, agg, na, bs, free=None):
    """Tentative prolongator from the rigid-body modes of every aggregate: level 0 rows [I | R(r)] (3 x 6, constrained
    dofs zeroed), higher levels [[I, R(r)], [0, I]] (6 x 6); r = node position minus aggregate centroid."""
    n = len(X)
    mem = agg >= 0
    cen = np.zeros((na, 3))
    cnt = np.bincount(agg[mem], minlength=na)
    for k in range(3):
        cen[:, k] = np.bincount(agg[mem], weights=X[mem, k], minlength=na) / np.maximum(cnt, 1)
    rows, cols, vals = [], [], []
    for i in np.nonzero(mem)[0]:
        r = X[i] - cen[agg[i]]
        P = np.zeros((bs, 6))
        P[0, 0] = P[1, 1] = P[2, 2] = 1.0
        P[0, 4], P[0, 5] = r[2], -r[1]
        P[1, 3], P[1, 5] = -r[2], r[0]
        P[2, 3], P[2, 4] = r[1], -r[0]
        if bs == 6:
            P[3, 3] = P[4, 4] = P[5, 5] = 1.0
        elif free is not None:
            P[~free[i]] = 0.0
        rr, cc = np.nonzero(P)
        rows.extend(i * bs + rr)
        cols.extend(agg[i] * 6 + cc)
        vals.extend(P[rr, cc])
    return sp.csr_matrix((vals, (rows, cols)), shape=(n * bs, na * 6)), cen


class Level:
    pass


def block_diag_inverse(A, bs):
    n = A.shape[0] // bs
    B = sp.bsr_matrix(A, blocksize=(bs, bs))
    B.sort_indices()
    D = np.zeros((n, bs, bs))
    for i in range(n):
        for t in range(B.indptr[i], B.indptr[i + 1]):
            if B.indices[t] == i:
                D[i] = B.data[t]
    D += 1e-8 * np.abs(D).max() * np.eye(bs)[None] if bs == 6 else 0.0
    Dinv = np.linalg.inv(D)
    return sp.bsr_matrix((Dinv, np.arange(n), np.arange(n + 1)), shape=A.shape).tocsr()


def build(A, X, free, part, keep_levels, max_levels=6):
    """keep_levels: number of coarse levels on which cross-part couplings are kept (0 = drop everywhere, big = keep)."""
    levels = []
    bs = 3
    excluded = ~free.any(axis=1)
    while True:
        L = Level()
        L.A, L.bs = A.tocsr(), bs
        L.Dinv = block_diag_inverse(L.A, bs)
        n = len(X)
        if n * bs <= 600 or len(levels) >= max_levels:
            L.dense = np.linalg.pinv(L.A.toarray())
            levels.append(L)
            break
        v = np.random.default_rng(0).standard_normal(n * bs)
        for _ in range(12):
            v = L.Dinv @ (L.A @ v)
            lam = np.linalg.norm(v)
            v /= lam
        L.lmax = 1.1 * lam
        G = sp.bsr_matrix(L.A, blocksize=(bs, bs))
        agg, na = aggregate(G.indptr, G.indices, n, excluded, part)
        L.P, cen = prolongator(X, agg, na, bs, free if bs == 3 else None)
        Ac = (L.P.T @ L.A @ L.P).tocsr()
        cpart = np.zeros(na, dtype=np.int64)
        cpart[agg[agg >= 0]] = part[agg >= 0]
        if len(levels) >= keep_levels:      # drop couplings between aggregates of different parts
            B = sp.bsr_matrix(Ac, blocksize=(6, 6)).tocoo()
            Ac = Ac.tocoo()
            same = cpart[Ac.row // 6] == cpart[Ac.col // 6]
            Ac = sp.csr_matrix((Ac.data[same], (Ac.row[same], Ac.col[same])), shape=Ac.shape)
        levels.append(L)
        A, X, part, bs = Ac, cen, cpart, 6
        excluded = np.zeros(na, dtype=bool)
    return levels


def cheb(L, b, x, degree, ratio=0.1):
    lmax, lmin = L.lmax, ratio * L.lmax
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma = theta / delta
    rho = 1.0 / sigma
    d = None
    for k in range(degree):
        r = b - L.A @ x if x is not None else b
        z = L.Dinv @ r
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
    if li == len(levels) - 1:
        return L.dense @ b
    x = cheb(L, b, None, 2)
    r = b - L.A @ x
    x = x + L.P @ vcycle(levels, li + 1, L.P.T @ r)
    # post-smoothing restarts the recurrence, as the library does
    lmax, lmin = L.lmax, 0.1 * L.lmax
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma = theta / delta
    rho = 1.0 / sigma
    d = None
    for k in range(2):
        z = L.Dinv @ (b - L.A @ x)
        if k == 0:
            d = z / theta
        else:
            rho_new = 1.0 / (2.0 * sigma - rho)
            d = rho_new * rho * d + 2.0 * rho_new / delta * z
            rho = rho_new
        x = x + d
    return x


def pcg_iterations(A, b, M, rtol=1e-10, maxit=400):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = r @ z
    bn = np.linalg.norm(b)
    for it in range(1, maxit + 1):
        Ap = A @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        if np.linalg.norm(r) <= rtol * bn:
            return it
        z = M(r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return maxit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=40)
    ap.add_argument("--parts", type=int, default=8)
    args = ap.parse_args()
    from glimslib_b200 import workloads as W, partition as P
    from oracle import fem
    w = W.c4_ellipsoid(args.grid)
    t = w["table"]
    prob = fem.Problem(w["mesh"].coords, w["mesh"].cells, w["cell_mat"],
                       fem.Materials(t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4]), w["dt"],
                       bc_dofs=w["bc_dofs"], bc_vals=w["bc_vals"])
    x = w["x0"].copy()
    F, J = fem.assemble(prob, x, x)
    nv = len(prob.coords)
    iu = (np.arange(prob.ndof) % 4) < 3
    K = J[iu][:, iu].tocsr()
    free = np.ones((nv, 3), dtype=bool)
    bc = prob.bc_dofs
    free[(bc // 4)[bc % 4 < 3], (bc % 4)[bc % 4 < 3]] = False
    keep = free.ravel().astype(float)
    K = (sp.diags(keep) @ K @ sp.diags(keep) + sp.diags(1.0 - keep)).tocsr()       # rows + columns, unit diagonal
    b = -(F[iu]) * keep
    out = {"grid": args.grid, "n_vertices": int(nv), "n_tets": int(len(prob.cells)), "parts": args.parts, "pcg_iterations": {}}
    variants = (("global", 1, 99), ("local-drop", args.parts, 0), ("local-keep1", args.parts, 1), ("local-keep", args.parts, 99))
    for name, parts, keep_levels in variants:
        t0 = time.time()
        part = P.rcb(prob.coords, parts).astype(np.int64)
        levels = build(K, prob.coords, free, part, keep_levels)
        its = pcg_iterations(K, b, lambda r: vcycle(levels, 0, r))
        out["pcg_iterations"][name] = its
        print("%-12s parts %d: %3d PCG iterations, levels %s (%.1f s)"
              % (name, parts, its, [L.A.shape[0] // L.bs for L in levels], time.time() - t0), file=sys.stderr)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
