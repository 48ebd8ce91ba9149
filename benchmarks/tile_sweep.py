#!/usr/bin/env python
"""Times the assembly kernels of config C4 on one GPU (CUDA events on the library stream, L2 flushed between
launches): the atomic / slice baselines and the fused tile kernel for several CTA sizes and column-split chunks.
Also checks the tile kernel's residual and matrices against the atomic kernel at full size.

    python benchmarks/tile_sweep.py [--grid 148] [--out gpurun_out/tile_sweep.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=148)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "tile_sweep.json"))
    ap.add_argument("--configs", default="128:12,128:24,128:8,256:12,256:24")
    args = ap.parse_args()
    from glimslib_b200 import workloads as W, _native as N
    sys.path.insert(0, ROOT)
    import bench
    t0 = time.time()
    w = W.c4_ellipsoid(args.grid)
    eng = W.build_engine(w)
    x = w["x0"].copy()
    rng = np.random.default_rng(0)
    x.reshape(-1, 4)[:, :3] = 1e-2 * rng.standard_normal((len(x) // 4, 3))
    eng.set_state(x)
    eng.set_prev(0.9 * x)
    print("setup %.1f s" % (time.time() - t0), file=sys.stderr)
    peak, _ = bench.peaks()
    ab = bench.algorithmic_bytes(3, eng.n_owned, eng.n_cells, eng.nnzb)
    out = {"grid": args.grid, "n_tets": int(eng.n_cells), "algorithmic_bytes": ab, "peak_gbs": peak, "runs": []}

    def rec(name, kid, variant, key):
        ms = eng.time_kernel(kid, variant, reps=10, flush_l2=True)
        r = {"name": name, "ms": ms, "gbs": ab[key] / ms / 1e6, "frac": ab[key] / ms / 1e6 / peak}
        out["runs"].append(r)
        print(json.dumps(r), file=sys.stderr)
        return r

    rec("full_atomic", 0, N.ASMK_ATOMIC, "assembly_full")
    rec("full_slice(+atomic residual)", 0, N.ASMK_SLICE, "assembly_full")
    rec("residual_atomic", 4, N.ASMK_ATOMIC, "residual")
    rec("residual+kcc_atomic", 5, N.ASMK_ATOMIC, "residual")
    # reference values for the parity check
    eng.assemble(what=N.ASM_ALL, kernel=N.ASMK_ATOMIC)
    F_ref = eng.residual()
    v = rng.standard_normal(eng.ndof)
    y_ref = eng.spmv(0, v)
    for cfg in args.configs.split(","):
        nt, ch = (int(t) for t in cfg.split(":"))
        t1 = time.time()
        eng.tile_config(nt, ch)
        eng.assemble(what=N.ASM_ALL, kernel=N.ASMK_TILE)
        tb = time.time() - t1
        info = eng.tile_info()
        F = eng.residual()
        y = eng.spmv(0, v)
        eF = float(np.abs(F - F_ref).max() / np.abs(F_ref).max())
        eK = float(np.abs(y - y_ref).max() / np.abs(y_ref).max())
        r = rec("full_tile nt=%d chunk=%d" % (nt, ch), 0, N.ASMK_TILE, "assembly_full")
        r.update(info=info, map_build_s=tb, err_F_vs_atomic=eF, err_Jv_vs_atomic=eK)
        rec("residual+kcc_tile nt=%d chunk=%d" % (nt, ch), 5, N.ASMK_TILE, "residual")
        rec("residual_tile nt=%d chunk=%d" % (nt, ch), 4, N.ASMK_TILE, "residual")
        print(json.dumps(r), file=sys.stderr)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out))
    eng.close()


if __name__ == "__main__":
    main()
