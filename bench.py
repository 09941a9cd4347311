#!/usr/bin/env python
"""Headline benchmark: HDG timesteps/s (BASELINE.json metric), Chorin projection k=2 on nx=1024.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nx NX] [--degree k]

A "step" is one Chorin timestep (`hdg_implicit.py:92-190`): BDM projection, tentative-velocity
solve, statically condensed mixed-Poisson solve (forward elimination, trace CG, back-substitution)
and the velocity/pressure update, on synthetic Taylor-Green data.

own arm (default)   device-resident timestepping through the Python timestepper mirror -> C-ABI;
                    `value` = timesteps/s with all inputs in HBM; `e2e` = the same step driven with
                    HOST buffers: per step the forcing comes from pinned host memory (H2D) and the
                    new velocity/pressure go back to the host (D2H), all inside the timed region.
reference arm       the CPU restatement of the reference path (oracle/, numpy+scipy sparse direct
                    solvers; the reference itself needs Firedrake/PETSc which cannot be installed
                    here) on a bounded sample, scaled to the same unit.

Prints ONE JSON line (see the task contract for the keys).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hdg_chorin_timesteps_per_second"
UNIT = "timesteps/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._thr = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 3 + j and r[3 + j].lower() == "active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's Chorin step on a bounded sample
# ------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    """one core's share: `nsteps` Chorin steps of the oracle on its own nx_sample mesh; returns (s/step, cells)"""
    nx_sample, degree, nsteps = job
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(1)
    except Exception:
        pass
    from incompressibleeulerhdg_b200.mesh import UnitSquareMesh
    from oracle.timesteppers import ChorinOracle, TaylorGreenOracle

    mesh = UnitSquareMesh(nx_sample, perturb=0.1)
    dt = 0.32 / nx_sample
    ts = ChorinOracle(mesh, degree, dt)
    prob = TaylorGreenOracle("exponential", 0.5)
    Q, p = ts.initial_state(prob)
    times = []
    for k in range(nsteps):
        t0 = time.perf_counter()
        Q, p = ts.step(Q, p, prob.f_rhs(k * dt))
        times.append(time.perf_counter() - t0)
    return float(np.mean(times)), mesh.nc


def cpu_chorin_sample(nx_sample, degree, nsteps, target_cells, cores=None):
    """The CPU restatement on all host cores: every core steps its own nx_sample mesh concurrently (the
    best case of a mesh-partitioned MPI run: no halo exchange, no load imbalance), and the aggregate
    cell-steps/s is scaled linearly to the target mesh (generous: sparse direct solvers grow faster)."""
    import multiprocessing as mp

    cores = max(1, min(os.cpu_count() or 1, 64) if cores is None else cores)
    if cores == 1:
        res = [_cpu_worker((nx_sample, degree, nsteps))]
    else:
        with mp.get_context("spawn").Pool(cores) as pool:
            res = pool.map(_cpu_worker, [(nx_sample, degree, nsteps)] * cores)
    sec_per_step = float(np.mean([r[0] for r in res]))
    cells = res[0][1]
    cell_steps_per_s = sum(r[1] / r[0] for r in res)
    return {
        "value": cell_steps_per_s / target_cells, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{nsteps} Chorin step(s) of oracle/timesteppers.py (numpy + scipy splu) on nx={nx_sample} k={degree} "
                  f"({cells} cells) on each of {cores} cores concurrently ({sec_per_step:.2f} s/step per core), "
                  f"aggregate cell-steps/s scaled linearly to {target_cells} cells",
    }, sec_per_step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nxm, nym, _ = mesh_shape(args, max(1, args.gpus))
    target_cells = 2 * nxm * nym
    t_all = []
    base = None
    for s in range(args.warmup + args.steps):
        b, sec = cpu_chorin_sample(args.cpu_nx, args.degree, 1, target_cells)
        if s >= args.warmup:
            t_all.append(1.0 / b["value"])
            base = b
    val = 1.0 / float(np.mean(t_all))
    base["value"] = val
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),
        "cpu_baseline": base,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "Firedrake/PETSc (the reference's own code path) is not installable in this image; this arm times "
                "the CPU restatement under oracle/",
    }
    print(json.dumps(line))


def mesh_shape(args, world):
    """(nx, ny, height): strong scaling partitions the named nx x nx unit-square mesh; weak scaling
    stacks `world` such squares into [0,1] x [0,world] so that every GPU keeps 2 nx^2 triangles"""
    if args.scaling == "weak":
        return args.nx, args.nx * world, float(world)
    return args.nx, args.nx, 1.0


def workload_config(args, world=1):
    nx, ny, height = mesh_shape(args, world)
    if world == 1:
        part = "single GPU"
    else:
        part = (f"{world} horizontal strips of {ny // world} rows of squares, one per GPU; vertex-adjacent ghost "
                "layer; NCCL halo exchange per neighbour-reading kernel, all-reduce per Krylov dot product")
    return {
        "workload": f"HDG Chorin projection (hdg_implicit.py, use_projection_method=True), k={args.degree}, "
                    f"{nx}x{ny} squares on [0,1]x[0,{height:g}] = {2 * nx * ny} triangles in total, upwind flux, "
                    f"Taylor-Green kappa=0.5, dt=0.32/nx, Krylov rtol {args.rtol:g}",
        "nx": nx, "ny": ny, "degree": args.degree, "dt": 0.32 / args.nx, "mesh_perturbation": 0.1,
        "cache": "inputs (>=1.1 GB trace matrix, 0.3 GB velocity fields per 2 nx^2 cells) exceed the 126 MB L2",
        "partition": part,
    }


# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # the version banner goes to stdout; this program prints ONE line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from incompressibleeulerhdg_b200.mesh import UnitSquareMesh
    from incompressibleeulerhdg_b200.model_problems import TaylorGreen
    from incompressibleeulerhdg_b200.timesteppers import IncompressibleEulerHDGImplicit

    nx, k = args.nx, args.degree
    dt = 0.32 / nx
    nxm, nym, height = mesh_shape(args, world)
    mesh = UnitSquareMesh(nxm, nym, perturb=0.1)
    if height != 1.0:  # weak scaling: [0,1] x [0,world]; the Taylor-Green field stays a no-flow solution there
        mesh.cell_xy[..., 1] *= height
    # torch.distributed is initialised => the timestepper partitions the mesh over the ranks
    ts = IncompressibleEulerHDGImplicit(mesh, k, dt, flux="upwind", use_projection_method=True, device=local,
                                        krylov_rtol=args.rtol)
    eng = ts.engine
    work_units = float(world) if args.scaling == "weak" else 1.0  # 2 nx^2-cell timesteps per global step
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q0, p0 = prob.initial_condition()
    f_rhs = prob.f_rhs()
    ts.initialise(Q0, p0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    step_no = 0
    # ---- device-resident arm -------------------------------------------------------------------
    for _ in range(args.warmup):
        ts.step(step_no, f_rhs)
        step_no += 1
    eng.reset_timers()
    l0 = eng.launch_count
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        ev0.record()
        for _ in range(args.steps):
            ts.step(step_no, f_rhs)
            step_no += 1
        ev1.record()
        barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = eng.launch_count - l0
    timers = eng.timers()
    its_p, its_t = ts.niter_pressure.value, ts.niter_tentative.value
    value = work_units * args.steps / (ms / 1e3)

    # ---- end-to-end arm: forcing from pinned host memory each step, (Q, p) back to the host ----
    from incompressibleeulerhdg_b200.functions import Function

    sQ, sp_, _ = eng.shapes()
    f_dev = Function(ts._V_Q)
    ts._V_Q.interpolate(f_rhs(0.0), out=f_dev)
    f_host = torch.from_numpy(eng.download(0, f_dev.data)).pin_memory()
    Q_host = torch.empty(sQ, dtype=torch.float64).pin_memory()
    p_host = torch.empty(sp_, dtype=torch.float64).pin_memory()

    def e2e_step(kstep):
        scale = float(np.exp(-0.5 * dt))  # host-side update of the forcing values for the next step
        eng.upload(0, f_host.numpy(), out=f_dev.data)
        ts.step(kstep, f_rhs, f_field=f_dev)
        eng.download(0, ts.Q.data, out=Q_host.numpy())
        eng.download(1, ts.p.data, out=p_host.numpy())
        return scale

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_step(step_no)
    step_no += 1
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step(step_no)
        step_no += 1
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = work_units * e2e_steps / e2e_s

    # ---- roofline of the dominant trace-solve kernel: back-to-back launches between two CUDA events on
    # the engine stream (the in-loop samples below also contain the host enqueue gap after the
    # per-iteration convergence check, so they under-report the kernel)
    xs, ys = eng.empty(2).normal_(), eng.empty(2)
    barrier()  # the e2e phase leaves the ranks skewed; the exchange inside the SpMV would wait for the slowest
    for _ in range(3):
        eng.trace_spmv_dev(xs, ys)
    barrier()
    n_spmv = 20
    sa, sb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sa.record()
    for _ in range(n_spmv):
        eng.trace_spmv_dev(xs, ys)
    sb.record()
    torch.cuda.synchronize()
    spmv_b2b_ms = sa.elapsed_time(sb) / n_spmv
    # the matrix-free advection operator of the tentative solve, timed the same way
    Xf, Yf = eng.empty(0).normal_(), eng.empty(0)
    for _ in range(2):
        eng.fimpl_apply_dev(ts._Q_star.data, Xf, Yf, c0=1.0, c1=-dt)
    barrier()
    n_fimpl = 10
    sa.record()
    for _ in range(n_fimpl):
        eng.fimpl_apply_dev(ts._Q_star.data, Xf, Yf, c0=1.0, c1=-dt)
    sb.record()
    torch.cuda.synchronize()
    fimpl_b2b_ms = sa.elapsed_time(sb) / n_fimpl
    del Xf, Yf
    # the kernel with the largest share of the step (profiles/launches_r1k.md: 39 %): one Chebyshev / facet-block-
    # Jacobi sweep on the facet Schur complement of the tentative-velocity preconditioner, timed the same way
    n_sweep, sweep_b2b_ms, sweep_err = 20, None, None
    try:
        sweep_b2b_ms = eng.tent_sweep_probe(dt, n_sweep)
    except Exception as exc:  # a measurement aid must not cost the bench line
        sweep_err = f"{type(exc).__name__}: {exc}"
    fp64_peak = eng.measure_fp64_peak()  # TFLOP/s, 8 independent DFMA chains per thread
    if world > 1:
        barrier()

    if rank == 0:
        peak, peak_src = load_peaks()
        b = k + 1
        nf_loc = eng.nf  # facets this rank's SpMV runs over (owned + ghost)
        nblocks = 5 * nf_loc
        spmv_bytes = nblocks * b * b * 8 + nblocks * 4 + 2 * b * nf_loc * 8
        spmv_ms, spmv_n = timers["spmv_sampled"]
        ach = spmv_bytes / spmv_b2b_ms / 1e6
        roofline = {
            "kernel": f"k_cg_spmv<{b}> (blocked-ELL trace SpMV of the CG; with N > 1 the launch includes this "
                      "rank's NCCL facet-halo exchange)",
            "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel at
            # nx=1024, k=2 on one GPU (profiles/ncu_r1b_cg_spmv_raw.csv.gz): 1.279 GB + 74.4 MB
            "traffic": 1353474384 if (world == 1 and nx == 1024 and k == 2) else None,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": spmv_bytes,
            "launch_ms": spmv_b2b_ms, "launches_timed": n_spmv,
            "in_loop_sampled_ms": spmv_ms / max(spmv_n, 1), "in_loop_samples": int(spmv_n),
        }
        # second-hottest kernel family (tentative velocity): k_fimpl is FP64-issue bound, not HBM bound
        dfma_per_cell = {1: 7 * 48 + 3 * 220, 2: 16 * 80 + 3 * 350, 3: 36 * 120 + 3 * 735, 4: 64 * 168 + 3 * 1176}[k]
        t = fimpl_b2b_ms
        other = {"k_fimpl": {
            "launch_ms": t, "bound": "fp64", "dfma_per_cell": dfma_per_cell,
            "achieved_tflops": 2.0 * dfma_per_cell * eng.nc / t / 1e9, "peak_tflops_measured": fp64_peak,
            "frac": 2.0 * dfma_per_cell * eng.nc / t / 1e9 / fp64_peak, "launches_timed": n_fimpl}}
        # per facet: geometry 6 doubles + 7 ints, and rhs, x, d (read), d, xout (written) of NM = k + 2 doubles each;
        # the 4 neighbour facets' x are re-reads of the same vector (L2)
        nm = k + 2
        sweep_bytes = (6 * 8 + 7 * 4 + 5 * nm * 8) * nf_loc
        if sweep_b2b_ms:
            sw = sweep_bytes / sweep_b2b_ms / 1e6
            other["k_tent_sweep"] = {
                "launch_ms": sweep_b2b_ms, "bound": "hbm", "algorithmic_bytes_per_launch": sweep_bytes,
                "achieved": sw, "peak": peak, "unit": "GB/s", "frac": sw / peak, "launches_timed": n_sweep,
                "share_of_step": "39 % of the device time of a step in profiles/launches_r1k.md"}
        else:
            other["k_tent_sweep"] = {"error": sweep_err}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu, _ = cpu_chorin_sample(args.cpu_nx, k, 1, mesh.nc)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {**workload_config(args, world), "tuning": dict(getattr(eng, "tuning", {}))},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(np.prod(sQ)) * 8,
                    "d2h_bytes_per_step": (int(np.prod(sQ)) + int(np.prod(sp_))) * 8, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "cuda_graph_replays": int(eng.graph_replays),
            "warm_start_restarts": int(eng.guess_restarts),
            "roofline": roofline,
            "other_kernels": other,
            "cpu_baseline": cpu,
            "iterations": {"trace_cg_per_solve": its_p, "tentative_bicgstab_per_solve": its_t,
                           "per_step_tentative_pressure": ts.iteration_history},
            "comm": {**eng.comm_stats(), "transport": ("nvlink-p2p" if getattr(eng, "p2p", False) else "nccl")
                     if world > 1 else "none", "p2p_timeouts": eng.p2p_status()},
            "breakdown_ms_per_step": {lab: timers[lab][0] / args.steps for lab in
                                      ("bdm_projection", "tentative_velocity_solve", "forward_elimination",
                                       "trace_solve", "back_substitution")},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=6,
                    help="untimed steps; the first ~6 steps after the interpolated initial condition are transient "
                         "(the discrete solution adjusts and the time-extrapolated Krylov guesses have no history yet)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=1024)
    ap.add_argument("--degree", type=int, default=2)
    ap.add_argument("--rtol", type=float, default=1e-12)
    ap.add_argument("--cpu-nx", type=int, default=16, help="mesh size of the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the named nx x nx mesh partitioned over the GPUs (BASELINE.json configs[2]); "
                         "weak: nx x (nx*gpus) squares, 2 nx^2 triangles per GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
