"""Time of one fused smoother step on level 1 of the V-cycle (glims_time_kernel 10), of the coarse correction below it (9) and
the K_uu iterations of two steps at C4 on one GPU; GLIMS_AMG_SORT=<window> sets the degree-sort window of the aggregates."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glimslib_b200 import workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
w = W.c4_ellipsoid(n)
eng = W.build_engine(w)
eng.set_prev(w["x0"])
eng.set_state(np.zeros_like(w["x0"]))
st = eng.step(4)
out = {"sort": os.environ.get("GLIMS_AMG_SORT", "default"), "grid": n, "its_u": [int(s["krylov_its_u"]) for s in st] if isinstance(st, list) else st}
for kid, name in ((10, "level1_step_us"), (9, "coarse_us"), (6, "fine_step_us")):
    out[name] = 1e3 * min(eng.time_kernel(kid, 0, reps=20, flush_l2=True) for _ in range(3))
print(json.dumps(out))
eng.close()
