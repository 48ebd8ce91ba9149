"""Multi-GPU driver: one process per GPU (torch.distributed for rendezvous only), one ``Engine`` per rank on its
share of the mesh.  The data path's collectives -- ghost halo exchange and the Krylov allreduce -- are issued by
the CUDA library itself over its own NCCL communicator (csrc/comm.cu)."""
import numpy as np

from . import partition as P
from .engine import Engine


def build_distributed_engine(w, rank, world, local_rank, dist=None, part=None):
    """Partition workload ``w`` (see workloads.py), create this rank's engine, wire the halo plan and NCCL."""
    mesh = w["mesh"]
    if part is None:
        part = P.rcb(mesh.coords, world)
    lm = P.build_local_mesh(mesh.coords, mesh.cells, part, rank, world)
    eng = Engine(lm.coords, lm.cells, w["cell_mat"][lm.cell_ids], device=local_rank, n_owned=lm.n_owned)
    eng.set_materials(w["table"])
    eng.set_dt(w["dt"])
    dofs, vals = lm.local_dofs(w["bc_dofs"], w["bc_vals"])
    eng.set_dirichlet(dofs, vals)
    eng.set_halo(lm.peers, lm.send_ptr, lm.send_idx, lm.recv_ptr)
    if dist is not None and world > 1:
        box = [Engine.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        eng.comm_init(world, rank, box[0])
    return eng, lm


def gather_owned(lm, x_local, dist, nb=None):
    """All ranks' owned values assembled into the global vector (on every rank)."""
    nb = lm.dim + 1 if nb is None else nb
    mine = (lm.l2g[:lm.n_owned], np.asarray(x_local).reshape(-1, nb)[:lm.n_owned])
    parts = [None] * lm.n_ranks
    dist.all_gather_object(parts, mine)
    out = np.zeros((int(lm._n_global), nb))
    for ids, vals in parts:
        out[ids] = vals
    return out.ravel()
