"""CPU tests of the multi-GPU host logic: RCB partition, one-layer overlap, halo plan -- including a
world_size-2 gloo run that performs the ghost exchange and the dot-product allreduce the CUDA library does
over NCCL, checked against the serial oracle operator."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from glimslib_b200 import partition as P
from glimslib_b200 import mesh as M
from oracle import fem


def _problem(n=6):
    m = M.box_mesh((0, 0, 0), (1, 1, 1), n, n, n)
    rng = np.random.default_rng(0)
    cm = rng.integers(0, 2, m.num_cells()).astype(np.int32)
    mats = fem.Materials.from_E_nu([3e-3, 1e-3], [0.45, 0.3], [0.1, 0.02], [0.2, 0.05], [0.15, 0.1])
    return m, fem.Problem(m.coords, m.cells, cm, mats, dt=0.5)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 8])
def test_rcb_is_balanced_and_complete(k):
    m, _ = _problem()
    part = P.rcb(m.coords, k)
    cnt = np.bincount(part, minlength=k)
    assert cnt.sum() == m.num_vertices() and cnt.max() - cnt.min() <= k
    assert np.array_equal(part, P.rcb(m.coords, k))       # deterministic


@pytest.mark.parametrize("k", [2, 4])
def test_local_meshes_cover_every_owned_row_and_plans_match(k):
    m, prob = _problem()
    part = P.rcb(m.coords, k)
    lms = [P.build_local_mesh(m.coords, m.cells, part, r, k) for r in range(k)]
    # every cell touching an owned vertex is present on that rank
    for lm in lms:
        touching = (part[m.cells] == lm.rank).any(axis=1)
        assert np.array_equal(np.nonzero(touching)[0], lm.cell_ids)
        assert np.array_equal(lm.l2g[lm.cells], m.cells[lm.cell_ids])
        assert np.all(part[lm.l2g[:lm.n_owned]] == lm.rank) and np.all(part[lm.l2g[lm.n_owned:]] != lm.rank)
    # what p sends to q is exactly what q expects from p, in the same order
    for p in lms:
        for ip, q in enumerate(p.peers):
            lq = lms[q]
            iq = list(lq.peers).index(p.rank)
            sent = p.l2g[p.send_idx[p.send_ptr[ip]:p.send_ptr[ip + 1]]]
            expected = lq.l2g[lq.n_owned + lq.recv_ptr[iq]: lq.n_owned + lq.recv_ptr[iq + 1]]
            assert np.array_equal(sent, expected)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m, prob = _problem()
        part = P.rcb(m.coords, world)
        lm = P.build_local_mesh(m.coords, m.cells, part, rank, world)
        nb = 4
        rng = np.random.default_rng(1)
        xg = rng.standard_normal(prob.ndof)
        # local operator: rows of owned vertices assembled from the local cells only (no assembly communication)
        lp = fem.Problem(lm.coords, lm.cells, prob.cell_mat[lm.cell_ids], prob.mats, prob.dt)
        xl = lm.to_local(xg)
        _, Jl = fem.assemble(lp, xl, xl)
        rows = slice(0, lm.n_owned * nb)
        # start from owned values only, then do the halo exchange (what comm.cu does with ncclSend/ncclRecv)
        v = rng.standard_normal(prob.ndof)
        vl = lm.to_local(v).reshape(-1, nb)
        vl[lm.n_owned:] = 0.0
        reqs, recv_bufs = [], []
        for ip, q in enumerate(lm.peers):
            sb = torch.from_numpy(np.ascontiguousarray(vl[lm.send_idx[lm.send_ptr[ip]:lm.send_ptr[ip + 1]]]))
            rb = torch.empty((int(lm.recv_ptr[ip + 1] - lm.recv_ptr[ip]), nb), dtype=torch.float64)
            recv_bufs.append((ip, rb))
            reqs.append(dist.isend(sb, int(q)))
            reqs.append(dist.irecv(rb, int(q)))
        for r in reqs:
            r.wait()
        for ip, rb in recv_bufs:
            vl[lm.n_owned + lm.recv_ptr[ip]: lm.n_owned + lm.recv_ptr[ip + 1]] = rb.numpy()
        assert np.array_equal(vl.ravel(), lm.to_local(v))
        yl = (Jl @ vl.ravel())[rows]
        # serial reference
        _, Jg = fem.assemble(prob, xg, xg)
        yg = (Jg @ v).reshape(-1, nb)[lm.l2g[:lm.n_owned]].ravel()
        err = np.abs(yl - yg).max() / np.abs(yg).max()
        # Krylov dot product: local owned part + allreduce
        t = torch.tensor([float(yl @ yl)], dtype=torch.float64)
        dist.all_reduce(t)
        dot_err = abs(t.item() - float((Jg @ v) @ (Jg @ v))) / float((Jg @ v) @ (Jg @ v))
        ret[rank] = (err, dot_err)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_halo_exchange_and_allreduce_match_serial():
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    for r in range(world):
        err, dot_err = ret[r]
        assert err < 1e-13 and dot_err < 1e-13
