"""Time of the coarse-grid correction below level 1 of the V-cycle (glims_time_kernel 9) at C4 on one GPU:
the fused persistent kernel (default) or, with GLIMS_AMG_FUSED=0, the launch sequence it replaces."""
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
eng.step(2)
out = {"fused": os.environ.get("GLIMS_AMG_FUSED", "1"), "grid": n}
for flush in (1, 0):
    out["us_flush%d" % flush] = 1e3 * min(eng.time_kernel(9, 0, reps=20, flush_l2=bool(flush)) for _ in range(3))
print(json.dumps(out))
eng.close()
