#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations (C1, C2, C3; C4 is bench.py's headline) on one GPU.
Prints one JSON line per config: steps/s over K steps after W warm-up steps, Newton / Krylov counts."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--configs", default="C1,C2,C3")
    args = ap.parse_args()
    import torch
    from glimslib_b200 import workloads as W
    builders = {"C1": W.c1_2d_subdomains, "C2": W.c2_2d_1m, "C3": W.c3_box}
    for name in args.configs.split(","):
        w = builders[name]()
        t0 = time.perf_counter()
        eng = W.build_engine(w)
        eng.set_prev(w["x0"])
        eng.set_state(np.zeros_like(w["x0"]))
        eng.step(args.warmup)
        torch.cuda.synchronize()
        setup = time.perf_counter() - t0
        t0 = time.perf_counter()
        st = eng.step(args.steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({"config": w["name"], "n_cells": int(w["mesh"].num_cells()), "n_dofs": int(eng.ndof),
                          "steps_per_s": args.steps / dt, "ms_per_step": 1e3 * dt / args.steps, "setup_s": setup,
                          "newton_its": float(np.mean([s["newton_its"] for s in st])),
                          "krylov_its_u": float(np.mean([s["krylov_its_u"] for s in st])),
                          "krylov_its_c": float(np.mean([s["krylov_its_c"] for s in st])),
                          "fnorm_last": st[-1]["fnorm"]}))
        eng.close()


if __name__ == "__main__":
    main()
