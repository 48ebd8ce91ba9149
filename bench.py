#!/usr/bin/env python
"""Headline benchmark: time steps per second of the mechanically-coupled reaction-diffusion model on the
3D 10M-tet brain-like ellipsoid (BASELINE.json metric), plus the HBM
roofline of the dominant kernels and a CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # one JSON line
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference path

A "step" is one backward-Euler step = one `solver.solve()` of the reference (simulation_base.py:302):
Newton-Krylov on the coupled system until the monolithic residual meets SNES rtol 1e-9 / atol 1e-10.
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

METRIC = "timesteps/sec, 3D 10M-tet coupled RD-mechanics (config C4)"
UNIT = "timesteps/s"


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
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
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


def oracle_problem(w):
    from oracle import fem
    t = w["table"]
    return fem.Problem(w["mesh"].coords, w["mesh"].cells, w["cell_mat"],
                       fem.Materials(t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4]), w["dt"],
                       bc_dofs=w["bc_dofs"], bc_vals=w["bc_vals"])


def _oracle_worker(job):
    """One host core: `warm` untimed + `steps` timed backward-Euler steps of the oracle on the C4 sample."""
    n_sample, warm, steps = job
    from glimslib_b200 import workloads as W
    from oracle import fem, solver as osolver
    w = W.c4_ellipsoid(n_sample)
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


def oracle_throughput(n_sample, warm, steps):
    """The CPU restatement on ALL host cores: one independent copy of the sample problem per core, each in its own
    process (`bench.py --cpu-worker`; scipy's sparse kernels are single-threaded, so this is how the port uses the machine).
    It ignores the communication a partitioned mpirun job would add, i.e. it flatters the CPU.  Returns (sample steps/s
    summed over the cores, cores, cells of the sample, seconds from the first start to the last finish)."""
    cores = max(1, min(os.cpu_count() or 1, 32))
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-worker", str(n_sample), str(warm), str(steps)]
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


def cpu_baseline(full_cells, steps=1, n_sample=24):
    """Oracle (numpy/scipy restatement of the reference path; Newton + GMRES(30)/ILU like PETSc's defaults) on a bounded
    sample: the same C4 configuration at a coarser voxel grid, one copy per host core.  Throughput is scaled linearly in
    the cell count to the full workload (optimistic for the CPU: its solve is superlinear)."""
    try:
        sps, cores, nc, dt = oracle_throughput(n_sample, 0, steps)
    except Exception as exc:      # e.g. a sandbox that forbids spawning: time one copy in this process instead
        print("cpu_baseline: process pool failed (%r), timing one core in-process" % (exc,), file=sys.stderr)
        t0, t1, nc = _oracle_worker((n_sample, 0, steps))
        sps, cores, dt = steps / (t1 - t0), 1, t1 - t0
    return {"value": sps * nc / full_cells, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "C4 at voxel grid n=%d (%d tets): %d step(s) on each of %d host cores in %.1f s = %.4f sample-steps/s "
                      "in total; scaled x(%d/%d) to the full mesh; scipy GMRES(30)+spilu, one process per core; "
                      "restatement, not FEniCS" % (n_sample, nc, steps, cores, dt, sps, nc, full_cells)}


def run_reference(args):
    """CPU arm: the oracle port timed on all host cores (FEniCS itself is not installable here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    full_cells = 10185024
    K, Wm = args.steps, args.warmup
    # bounded: each step is one backward-Euler step on the n=24 sample (~10 s per core); cap the run at a few minutes
    K = max(1, min(K, 6))
    Wm = min(Wm, 1)
    sps, cores, nc, dt = oracle_throughput(24, Wm, K)
    val = sps * nc / full_cells
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": Wm, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C4 3D voxel ellipsoid n=148, 10185024 tets, three tissues, coupled "
                                   "(timed on a bounded sample, see cpu_baseline.sample)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "C4 at n=24 (%d tets): %d steps on each of %d host cores in %.1f s, scaled x(%d/%d) to "
                                       "the full mesh; oracle port (scipy GMRES(30)+ILU, one process per core), FEniCS not "
                                       "installable offline" % (nc, K, cores, dt, nc, full_cells)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def algorithmic_bytes(d, n_v, n_c, nnzb):
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
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--grid", dest="n", type=int, default=148, help="voxel grid of the C4 ellipsoid (148 -> 10.19M tets)")
    ap.add_argument("--pc", default="amg", choices=["amg", "amg64", "jacobi"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--cpu-worker", nargs=3, type=int, default=None, help=argparse.SUPPRESS)   # n_sample warm steps
    args = ap.parse_args()
    if args.cpu_worker is not None:
        t0, t1, nc = _oracle_worker(tuple(args.cpu_worker))
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

    w = W.c4_ellipsoid(args.n)
    d = 3
    if world > 1:
        from glimslib_b200 import distributed as D
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        eng, part = D.build_distributed_engine(w, rank, world, local_rank, dist)
    else:
        eng = W.build_engine(w, device=local_rank)
        part = None
    opts = dict(pc={"amg": N.PC_AMG, "amg64": N.PC_AMG_FP64, "jacobi": N.PC_JACOBI}[args.pc])

    def local_vec(x):
        return x if part is None else part.to_local(x)

    x0 = local_vec(w["x0"])
    eng.set_prev(x0)
    eng.set_state(np.zeros_like(x0))

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: state already in HBM ------------------------------------------
    eng.step(W_steps, **opts)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = eng.launch_count
    t0 = time.perf_counter()
    stats = eng.step(K, **opts)
    barrier()
    t1 = time.perf_counter()
    launches = eng.launch_count - l0
    clocks = sampler.stop() if sampler else None
    elapsed = t1 - t0
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed = float(tt.item())
    dev_ms = float(np.mean([s["ms_total"] for s in stats]))

    # ---- end to end through the host-buffer API: H2D of the step inputs, D2H of the result -----
    ndof = eng.ndof
    pin_in = torch.empty(ndof, dtype=torch.float64).pin_memory().numpy()
    pin_out = torch.empty(ndof, dtype=torch.float64).pin_memory().numpy()
    pin_in[:] = eng.get_state()
    import ctypes
    lib = eng._lib
    K2 = max(3, min(K, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(K2):
        rc = lib.glims_set_prev(eng._h, N.as_dp(pin_in))          # H2D: u_previous
        rc |= lib.glims_set_state(eng._h, N.as_dp(pin_in))        # H2D: Newton start = last solution
        assert rc == 0
        eng.step(1, **opts)
        assert lib.glims_get_state(eng._h, N.as_dp(pin_out)) == 0   # D2H: the step's solution
        pin_in, pin_out = pin_out, pin_in                           # u_previous <- solution on the host (buffer swap)
    barrier()
    e2e_elapsed = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([e2e_elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_elapsed = float(tt.item())
    e2e = {"value": K2 / e2e_elapsed, "unit": UNIT, "h2d_bytes_per_step": 2 * ndof * 8 * world,
           "d2h_bytes_per_step": ndof * 8 * world, "steps": K2}

    # ---- roofline of the dominant kernels (CUDA events on the library's stream, L2 flushed) ----
    roof, kernels = None, {}
    if rank == 0 and not args.no_roofline:
        peak, peak_src = peaks()
        n_v, n_c, nnzb = eng.n_owned, eng.n_cells, eng.nnzb
        ab = algorithmic_bytes(d, n_v, n_c, nnzb)
        for name, kid, variant, key in (("smoother_step_fp16", 6, 0, "smoother_step"),
                                        ("spmv_kuu", 2, 0, "spmv_kuu"), ("spmv_mono", 1, 0, "spmv_mono"),
                                        ("spmv_kcc", 3, 0, "spmv_kcc"),
                                        ("assembly_full_tile", 0, 3, "assembly_full"),
                                        ("assembly_full_slice", 0, 2, "assembly_full"),
                                        ("assembly_full_gather", 0, 1, "assembly_full"),
                                        ("assembly_full_atomic", 0, 0, "assembly_full"),
                                        ("residual", 4, 0, "residual"), ("residual_kcc", 5, 0, "residual")):
            if kid == 6 and args.pc != "amg":
                continue
            # average of 10 launches, best of two such series (a stray series has been seen 1.5x slow)
            ms = min(eng.time_kernel(kid, variant, reps=10, flush_l2=True) for _ in range(2))
            gbs = ab[key] / ms / 1e6
            kernels[name] = {"ms": ms, "algorithmic_bytes": ab[key], "achieved_gbs": gbs, "frac": gbs / peak}
        one_gpu_default = (args.n == 148 and world == 1)
        if "smoother_step_fp16" in kernels:
            # dominant kernel of the step: 3 launches per PCG iteration, 25 % of the kernel time in the ncu launch list
            # of this command (profiles/r01_launch_summary_final.csv); next are the coarse-level smoother steps (22 %)
            # and the FP64 K_uu SpMV of PCG (18 %), both listed under "kernels"
            k = kernels["smoother_step_fp16"]
            roof = {"bound": "hbm", "kernel": "k_spmv32_row_cheb<3,__half> (fused Chebyshev smoother step, fine level of the V-cycle)",
                    "achieved": k["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": k["frac"],
                    # dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full of this kernel on this
                    # workload (profiles/r01_ncu_summary.md); only valid for the default mesh on one GPU
                    "traffic": 788.5e6 if one_gpu_default else None,
                    "traffic_source": "profiles/r01_ncu_summary.md (ncu --set full, r01 final build)", "peak_source": peak_src}
        else:
            k = kernels["spmv_kuu"]
            roof = {"bound": "hbm", "kernel": "k_spmv_block<3,3> (K_uu SELL-32 SpMV inside PCG)", "achieved": k["achieved_gbs"],
                    "peak": peak, "unit": "GB/s", "frac": k["frac"], "traffic": 2.070e9 if one_gpu_default else None,
                    "traffic_source": "profiles/r01_ncu_summary.md (ncu --set full, r01)", "peak_source": peak_src}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # reported at N=1 only
        cpu = cpu_baseline(w["mesh"].num_cells())

    if rank == 0:
        line = {
            "metric": METRIC, "value": K / elapsed, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_steps,
            "ms_per_step": 1e3 * elapsed / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "n_tets": int(w["mesh"].num_cells()),
                       "n_vertices": int(w["mesh"].num_vertices()), "n_dofs": int(w["mesh"].num_vertices() * 4),
                       "nnz_blocks": int(eng.nnzb), "dt": w["dt"], "solver": "block-triangular Newton-PCG, pc=%s (aggregation AMG; V-cycle in FP32 with an FP16 fine-level matrix, fused smoother steps) + successive-RHS projection; PCG iterations replayed as conditional CUDA graphs" % args.pc,
                       "tolerances": "SNES rtol 1e-9 atol 1e-10 (monolithic |F|), KSP rtol 1e-10",
                       "timing": "inputs larger than L2 (matrix 1.9 GB, vectors 42-56 MB); wall clock between "
                                 "device syncs, max over ranks",
                       "device_ms_per_step": dev_ms,
                       "newton_its_per_step": float(np.mean([s["newton_its"] for s in stats])),
                       "krylov_its_u_per_step": float(np.mean([s["krylov_its_u"] for s in stats])),
                       "krylov_its_c_per_step": float(np.mean([s["krylov_its_c"] for s in stats])),
                       "ms_assembly_per_step": float(np.mean([s["ms_assembly"] for s in stats])),
                       "ms_krylov_per_step": float(np.mean([s["ms_krylov"] for s in stats])),
                       "krylov_its_u_by_step": [int(s["krylov_its_u"]) for s in stats],
                       "final_fnorm": stats[-1]["fnorm"], "parallelism": "vertex partition x%d" % world},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "kernels": kernels,
            "cpu_baseline": cpu,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    eng.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
