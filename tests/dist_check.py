"""Multi-rank parity check, launched as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/dist_check.py
Runs the coupled model on a partitioned 3D box over N GPUs (NCCL halo exchange + allreduce inside the CUDA
library) and compares every step with a one-GPU run of the same library and with the CPU oracle."""
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
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = W.c3_box(12)
    eng, lm = D.build_distributed_engine(w, rank, world, local, dist)
    x0 = lm.to_local(w["x0"])
    eng.set_prev(x0)
    eng.set_state(np.zeros_like(x0))
    kw = dict(snes_rtol=1e-10, snes_atol=1e-14, ksp_rtol=1e-12)
    sols, its = [], []
    for _ in range(3):
        st = eng.step(1, **kw)[0]
        assert st["converged"] == 1
        its.append(st["krylov_its_u"])
        sols.append(D.gather_owned(lm, eng.get_state(), dist))
    if rank == 0:
        ref = W.build_engine(w, device=local)
        ref.set_prev(w["x0"])
        ref.set_state(np.zeros_like(w["x0"]))
        from oracle import fem, solver as osolver
        t = w["table"]
        prob = fem.Problem(w["mesh"].coords, w["mesh"].cells, w["cell_mat"],
                           fem.Materials(t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4]), w["dt"],
                           bc_dofs=w["bc_dofs"], bc_vals=w["bc_vals"])
        recs, _ = osolver.run(prob, w["x0"], 3, linear="gmres_ilu", rtol=1e-12, atol=1e-15, ksp_rtol=1e-13)
        its1 = []
        for k in range(3):
            its1.append(ref.step(1, **kw)[0]["krylov_its_u"])
            a, b, o = sols[k].reshape(-1, 4), ref.get_state().reshape(-1, 4), recs[k + 1][2].reshape(-1, 4)
            for name, sl in (("u", slice(0, 3)), ("c", slice(3, 4))):
                e1 = np.linalg.norm(a[:, sl] - b[:, sl]) / np.linalg.norm(b[:, sl])
                e2 = np.linalg.norm(a[:, sl] - o[:, sl]) / np.linalg.norm(o[:, sl])
                print("step %d %s: vs 1-GPU %.2e, vs oracle %.2e" % (k + 1, name, e1, e2))
                assert e1 < 1e-8 and e2 < 1e-7, (k, name, e1, e2)
        ref.close()
        print("K_uu PCG iterations per step: %d ranks %r, one GPU %r" % (world, its, its1))
        print("DIST_CHECK_OK world=%d" % world)
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
