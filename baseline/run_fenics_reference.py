#!/usr/bin/env python3
"""Run the UNMODIFIED reference (danielabler/glimslib on FEniCS 2017.2 / 2018.1) on one of the synthetic
configurations and dump what is needed to close the parity gap of DESIGN.md section 3.

This script cannot run in the build container (no FEniCS, no network).  On a host that has the reference's docker
image (``quay.io/dolfinadjoint/dolfin-adjoint``, dockerfiles/2017.2.0_libadjoint/Dockerfile:2) and the reference on
PYTHONPATH:

    mpirun -np 1 python3 baseline/run_fenics_reference.py --config C1 --out ref_C1.npz
    python3 baseline/compare_with_fenics.py ref_C1.npz           # on the B200 box

It records, per time step, ``sim.solution.vector()`` re-ordered to vertex-blocked order through
``vertex_to_dof_map`` (never the projected/saved fields, which carry ~1e-6 projection error), the
cell labels actually used, the initial vector, and wall-clock per step (for the FEniCS column of BASELINE.md).
Only serial runs dump vectors; under ``mpirun -np N`` only the timings are meaningful.
"""
import argparse
import time

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C1", choices=["C1", "C3"])
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--out", default="fenics_reference.npz")
    a = ap.parse_args()

    from glimslib import fenics_local as fenics                                  # the reference itself
    from glimslib.simulation.simulation_tumor_growth import TumorGrowth

    class Boundary(fenics.SubDomain):
        def inside(self, x, on_boundary):
            return on_boundary

    if a.config == "C1":
        n = a.n or 50
        mesh = fenics.RectangleMesh(fenics.Point(-5, -5), fenics.Point(5, 5), n, n)
        label_expr = fenics.Expression('(x[0]>=0.0) ? (1.0) : (2.0)', degree=1)
        ivc = fenics.Expression('sqrt(pow(x[0]-x0,2)+pow(x[1]-y0,2)) < 0.4 ? (1.0) : (0.0)', degree=1, x0=2.5, y0=2.5)
        ivu = fenics.Constant((0.0, 0.0))
        params = dict(diffusion={'outside': 0.0, 'A': 0.1, 'B': 0.0}, coupling={'outside': 0.0, 'A': 0.2, 'B': 0.0},
                      proliferation={'outside': 0.0, 'A': 0.1, 'B': 0.0}, E={'outside': 10E6, 'A': 0.001, 'B': 0.001},
                      poisson={'outside': 0.49, 'A': 0.40, 'B': 0.10})
        zero = fenics.Constant((0.0, 0.0))
    else:
        n = a.n or 55
        mesh = fenics.BoxMesh(fenics.Point(0, 0, 0), fenics.Point(1, 1, 1), n, n, n)
        label_expr = fenics.Expression('(x[0]<0.5) ? (1.0) : (2.0)', degree=0)
        ivc = fenics.Expression('exp(-60.0*(pow(x[0]-0.3,2)+pow(x[1]-0.5,2)+pow(x[2]-0.5,2)))', degree=1)
        ivu = fenics.Constant((0.0, 0.0, 0.0))
        params = dict(diffusion={'outside': 0.0, 'A': 2e-4, 'B': 0.0}, coupling={'outside': 0.0, 'A': 0.1, 'B': 0.0},
                      proliferation={'outside': 0.0, 'A': 0.05, 'B': 0.0}, E={'outside': 10E6, 'A': 3e-3, 'B': 1e-3},
                      poisson={'outside': 0.49, 'A': 0.45, 'B': 0.45})
        zero = fenics.Constant((0.0, 0.0, 0.0))
    labels = fenics.project(label_expr, fenics.FunctionSpace(mesh, "DG", 1))
    sim = TumorGrowth(mesh)
    sim.setup_global_parameters(label_function=labels, domain_names={0: 'outside', 1: 'A', 2: 'B'},
                                boundaries={'boundary_all': Boundary()},
                                dirichlet_bcs={'clamped': {'bc_value': zero, 'named_boundary': 'boundary_all', 'subspace_id': 0}},
                                von_neumann_bcs={})
    sim.setup_model_parameters(iv_expression={0: ivu, 1: ivc}, sim_time=a.steps, sim_time_step=1, **params)

    # replicate FenicsSimulation.run (simulation_base.py:236-317) with a dump after every solve
    W = sim.functionspace.function_space
    v2d = fenics.vertex_to_dof_map(W)                       # [vertex][component] -> dof, i.e. vertex-blocked order
    u_prev = sim.params.create_initial_value_function()
    sim._setup_problem(u_prev)
    out = {"x0": u_prev.vector().get_local()[v2d], "cell_labels": np.asarray(sim.subdomains.subdomains.array()),
           "coords": mesh.coordinates(), "cells": mesh.cells()}
    times = []
    for k in range(1, a.steps + 1):
        t0 = time.time()
        sim.solver.solve()
        times.append(time.time() - t0)
        out["x_%d" % k] = sim.solution.vector().get_local()[v2d]
        u_prev.assign(sim.solution)
    out["seconds_per_step"] = np.asarray(times)
    np.savez_compressed(a.out, **out)
    print("steps/s: %.4f (mean over %d steps, %d MPI ranks)" % (1.0 / np.mean(times), a.steps, fenics.MPI.size(mesh.mpi_comm())))


if __name__ == "__main__":
    main()
