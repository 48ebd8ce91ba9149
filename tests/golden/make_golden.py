#!/usr/bin/env python
"""Regenerates the fixtures in this directory from the CPU oracle (oracle/, Newton + sparse LU to rtol 1e-12).

These are REGRESSION anchors, not reference outputs: FEniCS cannot be imported in the build container, so there is no
reference-generated vector to commit ("parity unpinned").  They pin the oracle against accidental
change and give the GPU tests a fixed target that does not depend on re-running the oracle.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from glimslib_b200 import workloads as W  # noqa: E402
from oracle import fem, solver as osolver  # noqa: E402


def oracle_problem(w):
    t = w["table"]
    return fem.Problem(w["mesh"].coords, w["mesh"].cells, w["cell_mat"],
                       fem.Materials(t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4]), w["dt"],
                       bc_dofs=w["bc_dofs"], bc_vals=w["bc_vals"])


def cases():
    yield "c1_nx20_3steps", W.c1_2d_subdomains(nx=20), 3
    yield "c3_box6_2steps", W.c3_box(6), 2


def main():
    for name, w, steps in cases():
        recs, _ = osolver.run(oracle_problem(w), w["x0"], steps, linear="lu", rtol=1e-12, atol=1e-14)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), states=np.stack([r[2] for r in recs]),
                            times=np.array([r[0] for r in recs]))
        print(name, [float(np.linalg.norm(r[2])) for r in recs])


if __name__ == "__main__":
    main()
