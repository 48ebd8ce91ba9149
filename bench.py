#!/usr/bin/env python
"""Headline benchmark: time steps per second of the mechanically-coupled reaction-diffusion model on the
3D 10M-tet brain-like ellipsoid (BASELINE.json metric), plus the HBM
roofline of the kernels of the step and a CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # one JSON line
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference path
    python bench.py --workload C5 ...                        # 50M-tet mesh (BASELINE.json config 5); C3, C1, C2 likewise
    python bench.py --numbering random --reorder morton      # unstructured-ordering study

A "step" is one backward-Euler step = one `solver.solve()` of the reference (simulation_base.py:302):
Newton-Krylov on the coupled system until the monolithic residual meets SNES rtol 1e-9 / atol 1e-10.

Timed regions (all start from the same state: u_previous = initial condition, zero Newton start, empty projection basis):
  value     W warm-up steps, then K steps with the state resident in HBM (wall clock between device syncs, max over ranks)
  e2e       the same W + K step indices driven through the host-buffer C ABI: per step H2D of u_previous and of the
            Newton start from pinned memory, glims_step(1), D2H of the solution; the K steps after the W-th are timed
  value_cold / value_steady   steps 1-10 of the run / the last 10 timed steps, from the per-step device times
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "timesteps/s"
METRICS = {
    "C4": "timesteps/sec, 3D 10M-tet coupled RD-mechanics (config C4)",
    "C5": "timesteps/sec, 3D 50M-tet coupled RD-mechanics (config C5)",
    "C5W": "timesteps/sec, 3D 6.3M-tet-per-GPU coupled RD-mechanics (config C5, weak scaling)",
    "C3": "timesteps/sec, 3D 1M-tet box coupled RD-mechanics (config C3)",
    "C2": "timesteps/sec, 2D 1M-triangle reaction-diffusion (config C2)",
    "C1": "timesteps/sec, 2D 50x50 two-subdomain coupled case (config C1)",
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.split(",") for l in open(self.f.name).read().strip().splitlines() if l.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = [float(r[1]) for r in rows]
        out["sm_mhz"] = float(np.median(sm))
        out["sm_max_mhz"] = float(rows[0][2])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, nm in enumerate(names):
            if any("Active" in r[5 + i] and "Not" not in r[5 + i] for r in rows):
                out["reasons"].append(nm)
        out["samples"] = len(rows)
        return out


# ---------------------------------------------------------------------------------------------------------------
# workloads
def make_workload(name, grid, world=1, numbering="native", reorder="none"):
    from glimslib_b200 import workloads as W
    if name == "C4":
        w = W.c4_ellipsoid(grid or 148)
    elif name == "C5":
        w = W.c5_ellipsoid(grid or 252)
    elif name == "C5W":       # weak scaling: ~6.3M tets per GPU (n=252 on 8 GPUs)
        w = W.c5_ellipsoid(grid or int(round(252 * (world / 8.0) ** (1.0 / 3.0))))
    elif name == "C3":
        w = W.c3_box(grid or 55)
    elif name == "C2":
        w = W.c2_2d_1m(grid or 707)
    elif name == "C1":
        w = W.c1_2d_subdomains(grid or 50)
    else:
        raise SystemExit("unknown workload %r" % name)
    if numbering == "random":
        w = W.random_numbering(w, seed=0)
    if reorder == "morton":
        w = W.locality_numbering(w)
    return w


def oracle_problem(w):
    from oracle import fem
    t = w["table"]
    return fem.Problem(w["mesh"].coords, w["mesh"].cells, w["cell_mat"],
                       fem.Materials(t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4]), w["dt"],
                       bc_dofs=w["bc_dofs"], bc_vals=w["bc_vals"])


# ---------------------------------------------------------------------------------------------------------------
# CPU arms (the only place bench.py executes oracle/)
def _oracle_worker(job):
    """One host core: `warm` untimed + `steps` timed backward-Euler steps of the oracle on a workload sample."""
    name, n_sample, warm, steps = job
    from oracle import fem, solver as osolver
    w = make_workload(name, n_sample)
    prob = oracle_problem(w)
    geom = fem.geometry(prob.coords, prob.cells)
    x_prev = w["x0"].copy()
    x = np.zeros_like(x_prev)
    for _ in range(warm):
        x, _ = osolver.newton(prob, x, x_prev, linear="gmres_ilu", geom=geom)
        x_prev = x.copy()
    t0 = time.time()
    for _ in range(steps):
        x, _ = osolver.newton(prob, x, x_prev, linear="gmres_ilu", geom=geom)
        x_prev = x.copy()
    return t0, time.time(), int(w["mesh"].num_cells())


def oracle_throughput(name, n_sample, warm, steps, cores=None):
    """The CPU restatement on host cores: one independent copy of the sample problem per core, each in its own process
    (`bench.py --cpu-worker`; scipy's sparse kernels are single-threaded, so this is how the port uses the machine).  It
    ignores the communication a partitioned mpirun job would add, i.e. it flatters the CPU.  Returns (sample steps/s
    summed over the cores, cores, cells of the sample, seconds from the first start to the last finish)."""
    cores = cores or max(1, min(os.cpu_count() or 1, 32))
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-worker", name, str(n_sample), str(warm), str(steps)]
    procs = [subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for _ in range(cores)]
    res = []
    deadline = time.time() + 120 + 90 * (warm + steps)
    try:
        for p in procs:
            out, _ = p.communicate(timeout=max(1.0, deadline - time.time()))
            if p.returncode != 0:
                raise RuntimeError("cpu worker exited with %d" % p.returncode)
            res.append(json.loads(out.strip().splitlines()[-1]))
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    t0 = min(r["t0"] for r in res)
    t1 = max(r["t1"] for r in res)
    return cores * steps / (t1 - t0), cores, res[0]["cells"], t1 - t0


SAMPLE_GRID = {"C4": 24, "C5": 24, "C5W": 24, "C3": 20, "C2": 120, "C1": 50}


def cpu_baseline(name, full_cells, steps=1):
    """Oracle (numpy/scipy restatement of the reference path; Newton + GMRES(30)/ILU like PETSc's defaults) on a bounded
    sample: the same configuration at a coarser grid, one copy per host core.  Throughput is scaled linearly in the cell
    count to the full workload (optimistic for the CPU: its solve is superlinear).  EXTRAPOLATED, not measured at size --
    the measured same-size pair is `same_config`."""
    n_sample = SAMPLE_GRID[name]
    try:
        sps, cores, nc, dt = oracle_throughput(name, n_sample, 0, steps)
    except Exception as exc:      # e.g. a sandbox that forbids spawning: time one copy in this process instead
        print("cpu_baseline: process pool failed (%r), timing one core in-process" % (exc,), file=sys.stderr)
        t0, t1, nc = _oracle_worker((name, n_sample, 0, steps))
        sps, cores, dt = steps / (t1 - t0), 1, t1 - t0
    return {"value": sps * nc / full_cells, "unit": UNIT, "cores": cores, "kind": "port", "extrapolated": True,
            "sample": "%s at grid n=%d (%d cells): %d step(s) on each of %d host cores in %.1f s = %.4f sample-steps/s "
                      "in total; scaled x(%d/%d) to the full mesh; scipy GMRES(30)+spilu, one process per core; "
                      "restatement, not FEniCS" % (name, n_sample, nc, steps, cores, dt, sps, nc, full_cells)}


def same_config_pair(device, opts, grid=16, steps=2):
    """A MEASURED pair on identical inputs: config C3 (3D box, two tissues, coupled) at one grid size small enough for
    the CPU restatement to finish in the bench's time budget, the GPU library and the oracle both running the same
    `steps` backward-Euler steps from the same initial condition; the fields are compared as well."""
    from glimslib_b200 import workloads as W
    from oracle import fem, solver as osolver
    w = W.c3_box(grid)
    eng = W.build_engine(w, device=device)
    eng.prepare(**opts)
    eng.set_prev(w["x0"])
    eng.set_state(np.zeros_like(w["x0"]))
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.step(steps, **opts)
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    xg = eng.get_state().reshape(-1, 4)
    eng.close()
    prob = oracle_problem(w)
    geom = fem.geometry(prob.coords, prob.cells)
    x_prev = w["x0"].copy()
    x = np.zeros_like(x_prev)
    t0 = time.perf_counter()
    for _ in range(steps):
        x, _ = osolver.newton(prob, x, x_prev, linear="gmres_ilu", geom=geom)
        x_prev = x.copy()
    t_cpu = time.perf_counter() - t0
    xo = x.reshape(-1, 4)
    return {"workload": "C3 3D box %d^3 (%d tets, %d dofs), two tissues, coupled; %d steps from the initial condition"
                        % (grid, w["mesh"].num_cells(), len(w["x0"]), steps),
            "gpu": steps / t_gpu, "cpu": steps / t_cpu, "ratio": t_cpu / t_gpu, "unit": UNIT, "cpu_cores": 1,
            "cpu_kind": "oracle port: Newton + scipy GMRES(30)/spilu (PETSc-default-like), one process, scipy sparse "
                        "kernels are single-threaded",
            "gpu_includes": "steps only (setup excluded on both sides: oracle geometry / GPU pattern+AMG)",
            "rel_l2_u": float(np.linalg.norm(xg[:, :3] - xo[:, :3]) / np.linalg.norm(xo[:, :3])),
            "rel_l2_c": float(np.linalg.norm(xg[:, 3] - xo[:, 3]) / np.linalg.norm(xo[:, 3])),
            "note": "CPU arm at KSP rtol 1e-5 inside SNES rtol 1e-9 (PETSc defaults), GPU arm at KSP rtol 1e-10"}


def run_reference(args):
    """CPU arm: the oracle port timed on all host cores (FEniCS itself is not installable here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    n_sample = SAMPLE_GRID[name]
    full = make_workload(name, args.n, int(os.environ.get("WORLD_SIZE", "1")))
    full_cells = int(full["mesh"].num_cells())
    K, Wm = args.steps, args.warmup
    # bounded: each step is one backward-Euler step on the sample (~10 s per core); cap the run at a few minutes
    K = max(1, min(K, 6))
    Wm = min(Wm, 1)
    sps, cores, nc, dt = oracle_throughput(name, n_sample, Wm, K)
    val = sps * nc / full_cells
    line = {"impl": "reference", "metric": METRICS[name], "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": Wm, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": full["name"] + " (timed on a bounded sample, see cpu_baseline.sample)",
                       "n_tets": full_cells},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "extrapolated": True,
                             "sample": "%s at n=%d (%d cells): %d steps on each of %d host cores in %.1f s, scaled x(%d/%d) to "
                                       "the full mesh; oracle port (scipy GMRES(30)+ILU, one process per core), FEniCS not "
                                       "installable offline" % (name, n_sample, nc, K, cores, dt, nc, full_cells)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def algorithmic_bytes(d, n_v, n_c, nnzb):
    """Compulsory traffic per launch (DESIGN.md section 5 states each formula)."""
    nb = d + 1
    return {
        "spmv_kuu": (8 * d * d + 4) * nnzb + 16 * d * n_v,
        "spmv_mono": (8 * (d * d + d + 1) + 4) * nnzb + 16 * nb * n_v,
        "spmv_kcc": 12 * nnzb + 16 * n_v,
        "assembly_full": n_c * (4 * nb + 8 * d * nb + 4) + 8 * (d * d + d + 1) * nnzb + 3 * 8 * nb * n_v,
        "residual": n_c * (4 * nb + 8 * d * nb + 4) + 3 * 8 * nb * n_v,
        # fused Chebyshev smoother step of the V-cycle's fine level: FP16 d x d blocks + int32 column per block;
        # per vertex FP32 x (gathered, counted once), rhs, Dinv (d x d), d (read + write), own x, new x
        "smoother_step": (2 * d * d + 4) * nnzb + 4 * (d + d + d * d + 2 * d + d + d) * n_v,
        # row-walk K_cc + F_c: Klin read, column read, K_cc written per block; rho|K| + connectivity per cell;
        # c, M c_prev, f_ext,c read and F_c written per vertex
        "fc_kcc_rows": (8 + 4 + 8) * nnzb + (8 + 4 * nb) * n_c + 4 * 8 * n_v,
        # F_u = K_uu u + K_uc c: d x d and d x 1 blocks + column per block; x gathered (d+1), F_u written, f_ext,u read
        "fu_spmv": (8 * (d * d + d) + 4) * nnzb + 8 * (nb + 2 * d) * n_v,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="C4", choices=sorted(METRICS))
    ap.add_argument("--grid", dest="n", type=int, default=0, help="grid size of the workload (0: the named size)")
    ap.add_argument("--numbering", default="native", choices=["native", "random"],
                    help="random: vertices and cells randomly renumbered before the library sees them")
    ap.add_argument("--reorder", default="none", choices=["none", "morton"],
                    help="morton: Morton + degree-windowed renumbering applied on top (the library's locality reorder)")
    ap.add_argument("--pc", default="amg", choices=["amg", "amg64", "jacobi"])
    ap.add_argument("--asm", default="rows", choices=["rows", "atomic", "tile"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="N>1: skip the one-GPU reference run on rank 0")
    ap.add_argument("--cpu-worker", nargs=4, default=None, help=argparse.SUPPRESS)   # workload n_sample warm steps
    args = ap.parse_args()
    if args.cpu_worker is not None:
        nm, ns, wm, st = args.cpu_worker
        t0, t1, nc = _oracle_worker((nm, int(ns), int(wm), int(st)))
        print(json.dumps({"t0": t0, "t1": t1, "cells": nc}))
        return
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly one JSON line: everything libraries print (e.g. NCCL's version banner) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    from glimslib_b200 import _native as N
    from glimslib_b200 import workloads as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    W_steps = max(args.warmup, 3)
    K = args.steps

    t_setup = {}
    t0 = time.perf_counter()
    w = make_workload(args.workload, args.n, world, args.numbering, args.reorder)
    t_setup["mesh_host_s"] = time.perf_counter() - t0
    d = w["mesh"].dim
    nb = d + 1
    t0 = time.perf_counter()
    if world > 1:
        from glimslib_b200 import distributed as D
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        eng, part = D.build_distributed_engine(w, rank, world, local_rank, dist)
    else:
        dist = None
        eng = W.build_engine(w, device=local_rank)
        part = None
    torch.cuda.synchronize()
    t_setup["engine_s"] = time.perf_counter() - t0          # partition (N>1), H2D of the mesh, sparsity pattern, scatter map
    opts = dict(pc={"amg": N.PC_AMG, "amg64": N.PC_AMG_FP64, "jacobi": N.PC_JACOBI}[args.pc],
                asm_kernel={"rows": N.ASMK_ROWS, "atomic": N.ASMK_ATOMIC, "tile": N.ASMK_TILE}[args.asm])

    def local_vec(x):
        return x if part is None else part.to_local(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        tt = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    x0 = local_vec(w["x0"])
    eng.set_prev(x0)
    eng.set_state(np.zeros_like(x0))
    t0 = time.perf_counter()
    eng.prepare(**opts)           # K_uu / K_uc assembly + elimination, AMG hierarchy, row-walk maps
    barrier()
    t_setup["prepare_s"] = time.perf_counter() - t0
    setup_s = max_over_ranks(sum(t_setup.values()))

    def restart():
        eng.set_prev(x0)
        eng.set_state(np.zeros_like(x0))
        eng.reset_history()

    # ---- device-resident timing: state already in HBM ------------------------------------------
    restart()
    stats_w = eng.step(W_steps, **opts)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = eng.launch_count
    t0 = time.perf_counter()
    stats = eng.step(K, **opts)
    barrier()
    elapsed = max_over_ranks(time.perf_counter() - t0)
    launches = eng.launch_count - l0
    clocks = sampler.stop() if sampler else None
    all_stats = stats_w + stats
    ms_all = [s["ms_total"] for s in all_stats]
    n_cold = min(10, len(ms_all))
    n_steady = min(10, K)
    value_cold = n_cold / (sum(ms_all[:n_cold]) * 1e-3)
    value_steady = n_steady / (sum(ms_all[-n_steady:]) * 1e-3)
    dev_ms = float(np.mean([s["ms_total"] for s in stats]))

    # ---- the solution the timed run ended with (multi-rank parity evidence) ----------------------
    x_fin = eng.get_state().reshape(-1, nb)[:eng.n_owned]
    sums = torch.tensor([float((x_fin[:, :d] ** 2).sum()), float((x_fin[:, d] ** 2).sum()), float(x_fin.sum())],
                        device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(sums)
    solution = {"l2_u": float(sums[0].sqrt().item()), "l2_c": float(sums[1].sqrt().item()),
                "checksum": float(sums[2].item()), "after_steps": W_steps + K}
    x_glob = None
    if world > 1 and not args.no_verify:
        from glimslib_b200 import distributed as D
        x_glob = D.gather_owned(part, eng.get_state(), dist)

    # ---- end to end through the host-buffer API, same step indices: H2D of the step inputs, D2H of the result -----
    e2e = None
    ndof = eng.ndof
    if not args.no_e2e:
        pin_in = torch.empty(ndof, dtype=torch.float64).pin_memory().numpy()
        pin_new = torch.empty(ndof, dtype=torch.float64).pin_memory().numpy()
        pin_out = torch.empty(ndof, dtype=torch.float64).pin_memory().numpy()
        lib = eng._lib
        restart()
        pin_in[:] = x0                # u_previous
        pin_new[:] = 0.0              # Newton start of step 1 (stg:100); afterwards the last solution
        t_e2e0 = None
        for i in range(W_steps + K):
            if i == W_steps:
                barrier()
                t_e2e0 = time.perf_counter()
            rc = lib.glims_set_prev(eng._h, N.as_dp(pin_in))          # H2D: u_previous
            rc |= lib.glims_set_state(eng._h, N.as_dp(pin_new))       # H2D: Newton start
            assert rc == 0
            eng.step(1, **opts)
            assert lib.glims_get_state(eng._h, N.as_dp(pin_out)) == 0   # D2H: the step's solution
            pin_in, pin_out = pin_out, pin_in                           # u_previous <- solution (buffer swap on the host)
            pin_new = pin_in
        barrier()
        e2e_elapsed = max_over_ranks(time.perf_counter() - t_e2e0)
        e2e = {"value": K / e2e_elapsed, "unit": UNIT, "h2d_bytes_per_step": 2 * ndof * 8 * world,
               "d2h_bytes_per_step": ndof * 8 * world, "steps": K, "warmup": W_steps,
               "same_step_indices_as_value": True}

    # ---- roofline of the kernels of the step (CUDA events on the library's stream, L2 flushed) ----
    roof, kernels = None, {}
    if rank == 0 and not args.no_roofline:
        peak, peak_src = peaks()
        n_v, n_c, nnzb = eng.n_owned, eng.n_cells, eng.nnzb
        ab = algorithmic_bytes(d, n_v, n_c, nnzb)
        todo = [("smoother_step_fp16", 6, 0, "smoother_step"), ("spmv_kuu", 2, 0, "spmv_kuu"),
                ("fc_kcc_rows", 7, 0, "fc_kcc_rows"), ("fu_spmv", 8, 0, "fu_spmv"),
                ("coarse_vcycle_fused", 9, 0, None),
                ("spmv_kcc", 3, 0, "spmv_kcc"), ("spmv_mono", 1, 0, "spmv_mono"),
                ("residual_kcc", 5, 4, "residual"), ("residual_kcc_atomic", 5, 0, "residual"),
                ("assembly_full_tile", 0, 3, "assembly_full"), ("assembly_full_slice", 0, 2, "assembly_full"),
                ("assembly_full_atomic", 0, 0, "assembly_full"), ("residual_atomic", 4, 0, "residual")]
        for name, kid, variant, key in todo:
            if kid == 6 and args.pc != "amg":
                continue
            try:
                # average of 10 launches, best of two such series (a stray series has been seen 1.5x slow)
                ms = min(eng.time_kernel(kid, variant, reps=10, flush_l2=True) for _ in range(2))
            except Exception as exc:
                kernels[name] = {"error": str(exc)}
                continue
            if key is None:        # latency-bound: no byte model (levels >= 2 of the V-cycle, one persistent kernel)
                kernels[name] = {"ms": ms, "bound": "latency"}
                continue
            gbs = ab[key] / ms / 1e6
            kernels[name] = {"ms": ms, "algorithmic_bytes": ab[key], "achieved_gbs": gbs, "frac": gbs / peak}
        default_c4 = (args.workload == "C4" and not args.n and world == 1 and args.numbering == "native")
        kernels["_pattern"] = {"nnz_blocks": int(nnzb), "sell_slots": int(eng.nslots),
                               "sell_padding": float(eng.nslots) / float(nnzb) - 1.0}
        if "ms" in kernels.get("smoother_step_fp16", {}):
            # dominant kernel of the step (profiles/r02b_launch_summary.csv): 3 launches per PCG iteration
            k = kernels["smoother_step_fp16"]
            roof = {"bound": "hbm", "kernel": "k_spmv32_row_cheb<3,__half> (fused Chebyshev smoother step, fine level of the V-cycle)",
                    "achieved": k["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": k["frac"],
                    # dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full of this kernel on this
                    # workload (profiles/r02_ncu_summary.md); only valid for the default mesh on one GPU
                    "traffic": 746.7e6 if default_c4 else None,
                    "traffic_source": "profiles/r02_ncu_summary.md, second pass (ncu --set full, dram read 715.0 MB + write 31.8 MB)",
                    "peak_source": peak_src}
        else:
            k = kernels["spmv_kuu"]
            roof = {"bound": "hbm", "kernel": "k_spmv_block<3,3> (K_uu SELL-32 SpMV inside PCG)", "achieved": k["achieved_gbs"],
                    "peak": peak, "unit": "GB/s", "frac": k["frac"], "traffic": 2.070e9 if default_c4 else None,
                    "traffic_source": "profiles/r01_ncu_summary.md (ncu --set full, r01)", "peak_source": peak_src}

    # ---- N>1: the same run on ONE GPU (rank 0), compared pointwise with the gathered N-rank solution ----------
    parity = None
    if x_glob is not None and rank == 0:
        ref = W.build_engine(w, device=local_rank)
        ref.set_prev(w["x0"])
        ref.set_state(np.zeros_like(w["x0"]))
        ref.step(W_steps + K, **opts)
        a, b = x_glob.reshape(-1, nb), ref.get_state().reshape(-1, nb)
        ref.close()
        eu = float(np.linalg.norm(a[:, :d] - b[:, :d]) / np.linalg.norm(b[:, :d]))
        ec = float(np.linalg.norm(a[:, d] - b[:, d]) / np.linalg.norm(b[:, d]))
        parity = {"rel_l2_u": eu, "rel_l2_c": ec, "bar": 1e-8, "ok": bool(eu < 1e-8 and ec < 1e-8),
                  "l2_u_1gpu": float(np.linalg.norm(b[:, :d])), "l2_c_1gpu": float(np.linalg.norm(b[:, d])),
                  "what": "final state after %d steps, %d ranks (gathered) vs one GPU, same library and options" % (W_steps + K, world)}
    if world > 1:
        dist.barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # reported at N=1 only
        cpu = cpu_baseline(args.workload, w["mesh"].num_cells())
        if d == 3:
            try:
                cpu["same_config"] = same_config_pair(local_rank, opts)
            except Exception as exc:
                cpu["same_config"] = {"error": repr(exc)}

    if rank == 0:
        line = {
            "metric": METRICS[args.workload], "value": K / elapsed, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W_steps, "ms_per_step": 1e3 * elapsed / K, "higher_is_better": True,
            "scaling": "weak" if args.workload == "C5W" else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "value_cold": value_cold, "value_steady": value_steady, "setup_s": setup_s,
            "config": {"workload": w["name"], "n_tets": int(w["mesh"].num_cells()),
                       "n_vertices": int(w["mesh"].num_vertices()), "n_dofs": int(w["mesh"].num_vertices() * nb),
                       "nnz_blocks": int(eng.nnzb), "dt": w["dt"],
                       "solver": "block-triangular Newton-PCG, pc=%s (aggregation AMG; V-cycle in FP32 with an FP16 fine-level matrix, fused smoother steps) + successive-RHS projection; PCG iterations replayed as conditional CUDA graphs; per-Newton assembly=%s" % (args.pc, args.asm),
                       "tolerances": "SNES rtol 1e-9 atol 1e-10 (monolithic |F|), KSP rtol 1e-10 (glims_default_opts; "
                                     "tests/test_gpu_parity.py::test_default_tolerances_meet_the_parity_bar)",
                       "timing": "inputs larger than L2 (matrix 1.9 GB, vectors 42-56 MB); wall clock between "
                                 "device syncs, max over ranks; value/e2e both start from the initial condition with an "
                                 "empty projection basis: steps %d..%d are timed" % (W_steps + 1, W_steps + K),
                       "value_cold": "steps 1-%d (device time per step)" % n_cold,
                       "value_steady": "last %d timed steps (device time per step)" % n_steady,
                       "setup_breakdown_s": t_setup,
                       "device_ms_per_step": dev_ms,
                       "newton_its_per_step": float(np.mean([s["newton_its"] for s in stats])),
                       "krylov_its_u_per_step": float(np.mean([s["krylov_its_u"] for s in stats])),
                       "krylov_its_c_per_step": float(np.mean([s["krylov_its_c"] for s in stats])),
                       "ms_assembly_per_step": float(np.mean([s["ms_assembly"] for s in stats])),
                       "ms_krylov_per_step": float(np.mean([s["ms_krylov"] for s in stats])),
                       "krylov_its_u_by_step": [int(s["krylov_its_u"]) for s in all_stats],
                       "ms_by_step": [round(float(m), 3) for m in ms_all],
                       "launches_per_step": launches / float(K),
                       "final_fnorm": stats[-1]["fnorm"], "parallelism": "vertex partition x%d" % world},
            "solution": solution, "parity_vs_1gpu": parity,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "kernels": kernels,
            "cpu_baseline": cpu,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        raise SystemExit("bench.py: %d-rank solution differs from the one-GPU solution: %r" % (world, parity))


if __name__ == "__main__":
    main()
