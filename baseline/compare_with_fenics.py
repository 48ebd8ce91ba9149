#!/usr/bin/env python3
"""Compare a dump of baseline/run_fenics_reference.py with the B200 path on the same mesh, labels, initial vector
and time steps: prints the relative L2 difference of concentration and displacement per step (target 1e-8)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ref = np.load(sys.argv[1])
    from glimslib_b200 import workloads as W
    from glimslib_b200.engine import Engine
    coords, cells, lab = ref["coords"], ref["cells"].astype(np.int32), ref["cell_labels"]
    d = coords.shape[1]
    w = W.c1_2d_subdomains() if d == 2 else W.c3_box()
    assert np.allclose(w["mesh"].coords, coords) and np.array_equal(w["mesh"].cells, cells), "mesh numbering differs"
    labels, cell_mat = np.unique(lab, return_inverse=True)
    eng = Engine(coords, cells, cell_mat.astype(np.int32))
    eng.set_materials(w["table"][labels - 1])
    eng.set_dt(1.0)
    eng.set_dirichlet(w["bc_dofs"], w["bc_vals"])
    eng.set_prev(ref["x0"])
    eng.set_state(np.zeros_like(ref["x0"]))
    nb = d + 1
    k = 1
    while "x_%d" % k in ref:
        eng.step(1, snes_rtol=1e-11, snes_atol=1e-14, ksp_rtol=1e-12)
        x, r = eng.get_state().reshape(-1, nb), ref["x_%d" % k].reshape(-1, nb)
        ec = np.linalg.norm(x[:, d] - r[:, d]) / np.linalg.norm(r[:, d])
        eu = np.linalg.norm(x[:, :d] - r[:, :d]) / max(np.linalg.norm(r[:, :d]), 1e-300)
        print("step %3d  rel L2: concentration %.3e  displacement %.3e" % (k, ec, eu))
        k += 1
    print("FEniCS steps/s %.4f" % (1.0 / ref["seconds_per_step"].mean()))


if __name__ == "__main__":
    main()
