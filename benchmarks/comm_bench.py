#!/usr/bin/env python
"""Latency of the data-path collectives on N GPUs, peer-memory transport vs NCCL (CUDA events, back-to-back calls):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/comm_bench.py [--grid 60]
Prints one JSON line on rank 0 (max over ranks of the per-call time in microseconds)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from glimslib_b200 import distributed as D  # noqa: E402
from glimslib_b200 import workloads as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=60)
    ap.add_argument("--reps", type=int, default=500)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = W.c4_ellipsoid(args.grid)
    eng, lm = D.build_distributed_engine(w, rank, world, local, dist)
    out = {"n_gpus": world, "grid": args.grid, "n_owned": int(lm.n_owned), "n_ghost": int(eng.n_vertices - lm.n_owned),
           "n_peers": int(len(lm.peers))}
    names = {0: "halo_f64_x4", 1: "halo_f32_x3", 2: "allreduce_2"}
    for transport in ("p2p", "nccl"):
        in_use = eng.set_p2p(transport == "p2p")
        out[transport + "_in_use"] = in_use if transport == "p2p" else (not in_use)
        for kind, name in names.items():
            dist.barrier()
            us = eng.comm_bench(kind, args.reps)
            t = torch.tensor([us], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out["%s_%s_us" % (transport, name)] = round(float(t.item()), 2)
    eng.set_p2p(True)
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
