"""Cost of the device discrete adjoint (glims_adjoint) at C4 on one GPU: N forward steps alone against
N forward steps + misfit + backward sweep + per-material gradient, both from the same initial state."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glimslib_b200 import workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
w = W.c4_ellipsoid(n)
eng = W.build_engine(w)
nv, d = eng.n_vertices, eng.dim


def restart():
    eng.set_prev(w["x0"])
    eng.set_state(np.zeros_like(w["x0"]))
    eng.reset_history()


restart()
eng.step(2)                       # hierarchy, graphs
restart()
t0 = time.perf_counter()
st = eng.step(n_steps)
t_fwd = time.perf_counter() - t0
xN = eng.get_state().reshape(nv, d + 1)
levels = [0.2, 0.5]
# targets: the thresholded final state of this run, shifted, so that every misfit term and its derivative are non-zero
tg = np.stack([0.5 * (np.tanh((xN[:, d] * 0.9 - lv) / 0.01) + 1.0) for lv in levels])
ut = 0.9 * xN[:, :d]
restart()
t0 = time.perf_counter()
J, grad = eng.adjoint_gradient(n_steps, levels, tg, ut)
t_adj = time.perf_counter() - t0
print(json.dumps({"grid": n, "n_cells": int(w["mesh"].num_cells()), "n_steps": n_steps, "forward_s": t_fwd, "forward_plus_adjoint_s": t_adj,
                  "ratio": t_adj / t_fwd, "krylov_its_u_forward": int(sum(s["krylov_its_u"] for s in st)), "J": J,
                  "grad": grad.tolist()}))
eng.close()
