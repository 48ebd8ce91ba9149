"""A small coupled 3D run (box n^3, default 18: fine level + smoothed level 1 + dense coarsest level), two steps, for tools that
slow a run down by orders of magnitude (compute-sanitizer memcheck / racecheck with --kernel-name kns=k_split_tma ..., a debugger).
compute-sanitizer is closed on the pool this round was built on, so no sanitizer log is committed; shared-memory and barrier
discipline of the bulk-copy kernel is argued in its comments and exercised by the parity tests at several ring shapes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glimslib_b200 import workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = W.c3_box(n)
eng = W.build_engine(w)
eng.set_prev(w["x0"])
eng.set_state(np.zeros_like(w["x0"]))
st = eng.step(steps)
x = eng.get_state()
print("steps", steps, "newton", [s["newton_its"] for s in st], "its_u", [s["krylov_its_u"] for s in st],
      "its_c", [s["krylov_its_c"] for s in st], "|x|", float(np.linalg.norm(x)))
eng.close()
