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
reference arm       the compiled CPU restatement of the reference path (oracle/cpu_ref, C++/OpenMP on all host
                    cores; the reference itself needs Firedrake/PETSc, which cannot be installed here) on a
                    bounded sample (nx = 256, 1/16 of the cells), scaled to the same unit.

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
class CpuChorin:
    """The compiled CPU restatement (oracle/cpu_ref/hdg_cpu_ref.cpp: C++/OpenMP, all host cores) stepping the same
    workload -- Chorin k=2, dt = 0.32/nx, Taylor-Green, rtol as the GPU arm -- on an nx_sample x nx_sample mesh.  The
    result is scaled linearly in the number of cells to the target mesh (generous to the CPU: the iteration counts of its
    Krylov solvers do not drop on finer meshes).  Timer labels are the reference's (src/auxilliary/logging.py:11-31)."""

    def __init__(self, nx_sample, degree, rtol, target_cells):
        from incompressibleeulerhdg_b200.mesh import UnitSquareMesh
        from oracle.cpu_ref import ChorinCpuRef
        from oracle.timesteppers import TaylorGreenOracle

        self.nx, self.k, self.target_cells = nx_sample, degree, target_cells
        self.mesh = UnitSquareMesh(nx_sample, perturb=0.1)
        self.dt = 0.32 / nx_sample
        t0 = time.perf_counter()
        self.ref = ChorinCpuRef(self.mesh, degree, self.dt, rtol=rtol, maxit=2000)
        self.prob = TaylorGreenOracle("exponential", 0.5)
        self.Q, self.p = self.ref.initial_state(self.prob)
        self.setup_s = time.perf_counter() - t0
        self.nstep = 0
        self.factor = target_cells / self.mesh.nc

    def step(self):
        """one timestep including the interpolation of the forcing (hdg_implicit.py:100); returns seconds"""
        t0 = time.perf_counter()
        f = self.ref.forcing(self.prob.f_rhs(self.nstep * self.dt))
        self.ref.step(self.Q, self.p, f)
        self.nstep += 1
        return time.perf_counter() - t0

    def describe(self, sec_per_step, nsteps):
        its = self.ref.iterations[-nsteps:]
        tm = self.ref.timers()
        n = max(1, tm["timestep"]["calls"])
        return {
            "value": 1.0 / (sec_per_step * self.factor), "unit": UNIT, "cores": self.ref.threads, "kind": "port",
            "sample": f"{nsteps} Chorin step(s) of oracle/cpu_ref/hdg_cpu_ref.cpp (C++/OpenMP, {self.ref.threads} threads; "
                      f"BiCGStab + facet-multiplier preconditioner, two-level CG on the condensed trace system, rtol as the "
                      f"GPU arm) on nx={self.nx} k={self.k} ({self.mesh.nc} cells, dt=0.32/nx): {sec_per_step:.2f} s/step, "
                      f"scaled linearly in cells (x{self.factor:g}) to {self.target_cells} cells",
            "sample_seconds_per_step": sec_per_step, "extrapolation_factor": self.factor, "setup_seconds": self.setup_s,
            "iterations_tentative_pressure": [list(i) for i in its],
            "seconds_per_step_by_label": {lab: v["seconds"] / n for lab, v in tm.items()},
            "checked_against": "oracle/ (numpy, pinned to the reference's forms) to 1e-10: tests/test_cpu_ref.py",
        }


def cpu_baseline_sample(args, target_cells):
    """the CPU baseline of the own arm's line: one warm-up step and `--cpu-steps` timed steps of the compiled CPU port"""
    cpu = CpuChorin(args.cpu_nx, args.degree, args.rtol, target_cells)
    cpu.step()
    secs = [cpu.step() for _ in range(max(1, args.cpu_steps))]
    return cpu.describe(float(np.mean(secs)), len(secs))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nxm, nym, _ = mesh_shape(args, max(1, args.gpus))
    target_cells = 2 * nxm * nym
    cpu = CpuChorin(args.cpu_nx, args.degree, args.rtol, target_cells)
    for _ in range(args.warmup):
        cpu.step()
    secs = [cpu.step() for _ in range(args.steps)]
    base = cpu.describe(float(np.mean(secs)), len(secs))
    val = base["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),
        "cpu_baseline": base,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "timed_region_s": float(np.sum(secs)),
        "note": "Firedrake/PETSc (the reference's own code path) is not installable in this image; this arm times the "
                "compiled CPU restatement under oracle/cpu_ref on all host cores, each step = one timestep of the same "
                "workload on a mesh with 1/extrapolation_factor of the cells",
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
# dram__bytes_read.sum + dram__bytes_write.sum of one k_tent_sweep32 launch inside a step (ncu --set full), keyed by
# (fp32 storage, ranks, nx, k); filled in from the capture of the kernel version that is shipped
SWEEP_TRAFFIC = {}
SWEEP_TRAFFIC_SOURCE = ("no capture of a later sweep of the shipped kernel; profiles/r2/ncu_r2s_tent_sweep32_raw.csv.gz is a "
                        "first (zero-iterate) sweep: 395 MB in 64 us")


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
    from incompressibleeulerhdg_b200.functions import Function
    from incompressibleeulerhdg_b200.mesh import UnitSquareMesh
    from incompressibleeulerhdg_b200.model_problems import TaylorGreen
    from incompressibleeulerhdg_b200.timesteppers import IncompressibleEulerHDGImplicit

    nx, k = args.nx, args.degree
    dt = 0.32 / nx
    nxm, nym, height = mesh_shape(args, world)
    t_setup = time.perf_counter()
    mesh = UnitSquareMesh(nxm, nym, perturb=0.1)
    if height != 1.0:  # weak scaling: [0,1] x [0,world]; the Taylor-Green field stays a no-flow solution there
        mesh.cell_xy[..., 1] *= height
    # torch.distributed is initialised => the timestepper partitions the mesh over the ranks
    ts = IncompressibleEulerHDGImplicit(mesh, k, dt, flux="upwind", use_projection_method=True, device=local,
                                        krylov_rtol=args.rtol, warm_start=True, warm_order=args.warm_order)
    eng = ts.engine
    work_units = float(world) if args.scaling == "weak" else 1.0  # 2 nx^2-cell timesteps per global step
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q0, p0 = prob.initial_condition()
    f_rhs = prob.f_rhs()
    ts.initialise(Q0, p0)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup
    mem_gb = torch.cuda.mem_get_info(local)
    mem_used_gb = (mem_gb[1] - mem_gb[0]) / 2**30

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

    def l2(f):
        return float(np.sqrt(eng.l2_inner_dev(f.space.kind, f.data, f.data)))

    def errors_and_checksums(t_now):
        """L2 errors against model_problem.solution(t) (`driver.py:365-380`) and rank-reduced checksums of the fields:
        the 1- and N-GPU runs of the same config must print the same numbers (to solver tolerance)"""
        Qe, pe = prob.solution(t_now)
        dQ, dp_ = Function(ts._V_Q), Function(ts._V_p)
        eng.lincomb_dev(dQ.data, [(1.0, ts.Q.data), (-1.0, Qe.data)])
        eng.lincomb_dev(dp_.data, [(1.0, ts.p.data), (-1.0, pe.data)])
        return {"t": t_now, "l2_error_velocity": l2(dQ), "l2_error_pressure": l2(dp_),
                "l2_norm_velocity": l2(ts.Q), "l2_norm_pressure": l2(ts.p),
                "l2_inner_velocity_exact": float(eng.l2_inner_dev(0, ts.Q.data, Qe.data))}

    def timed_steps(nsteps, step_no, step_fn=None):
        """(ms max over ranks, launches, iteration summary) of `nsteps` steps bracketed by barrier + synchronize"""
        step_fn = step_fn or (lambda kk: ts.step(kk, f_rhs))
        h0 = len(ts.iteration_history)
        l0 = eng.launch_count
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(nsteps):
            step_fn(step_no)
            step_no += 1
        ev1.record()
        barrier()
        hist = ts.iteration_history[h0:]
        its = {"tentative_per_solve": float(np.mean([a for a, _ in hist])), "trace_cg_per_solve": float(np.mean([b for _, b in hist])),
               "per_step_tentative_pressure": hist}
        return max_over_ranks(ev0.elapsed_time(ev1)), eng.launch_count - l0, its, step_no

    step_no = 0
    # ---- device-resident arm (the headline `value`) ----------------------------------------------
    for _ in range(args.warmup):
        ts.step(step_no, f_rhs)
        step_no += 1
    eng.reset_timers()
    kc0 = eng.kernel_counts()
    st0 = eng.tentative_stats()
    mx0 = eng.mixed_stats()
    with ClockSampler(local) as clocks:
        ms, launches, its_main, step_no = timed_steps(args.steps, step_no)
    kc1 = eng.kernel_counts()
    st1 = eng.tentative_stats()
    mx1 = eng.mixed_stats()
    timers = eng.timers()
    value = work_units * args.steps / (ms / 1e3)
    check_main = errors_and_checksums(step_no * dt)
    per_step_launches = {kname: (kc1[kname] - kc0.get(kname, 0)) / args.steps for kname in kc1 if kc1[kname] != kc0.get(kname, 0)}

    # ---- end-to-end arm: the same number of steps driven with HOST buffers.  Per step the forcing field comes from
    # pinned host memory (H2D) and the new velocity and pressure go back to pinned host memory (D2H); the transfers run
    # on the engine's copy stream (hdg_upload_begin / hdg_download_begin) and overlap the solver kernels: the upload
    # of step n + 1 and the download of step n - 1 are in flight while step n computes.
    sQ, sp_, _ = eng.shapes()
    f_dev = Function(ts._V_Q)
    ts._V_Q.interpolate(f_rhs(step_no * dt), out=f_dev)
    f_host = [torch.from_numpy(eng.download(0, f_dev.data)).pin_memory() for _ in range(2)]
    Q_host = [torch.empty(sQ, dtype=torch.float64).pin_memory() for _ in range(2)]
    p_host = [torch.empty(sp_, dtype=torch.float64).pin_memory() for _ in range(2)]
    e2e_state = {"n": 0}
    eng.upload_begin(0, f_host[0].numpy(), slot=0)

    def e2e_step(kstep):
        n = e2e_state["n"]
        slot = n % 2
        eng.upload_end(0, f_dev.data, slot=slot)              # forcing of this step (started one step ago)
        # the forcing of the next step starts travelling now (its values are the caller's business: updating 42 M
        # doubles on the host is not part of the path, so the two host buffers keep the forcing of the first step)
        eng.upload_begin(0, f_host[1 - slot].numpy(), slot=1 - slot)
        ts.step(kstep, f_rhs, f_field=f_dev)
        eng.download_begin(0, ts.Q.data, Q_host[slot].numpy(), slot=slot)
        eng.download_begin(1, ts.p.data, p_host[slot].numpy(), slot=slot)
        e2e_state["n"] = n + 1

    e2e_steps = args.steps if args.e2e_steps <= 0 else min(args.steps, args.e2e_steps)
    e2e_step(step_no)
    step_no += 1
    eng.copy_wait()
    t0 = time.perf_counter()
    e2e_ms, _, its_e2e, step_no = timed_steps(e2e_steps, step_no, e2e_step)
    eng.copy_wait()  # the last step's results have arrived on the host
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = work_units * e2e_steps / e2e_s
    e2e_probe = float(Q_host[(e2e_state["n"] - 1) % 2].abs().max())  # the result was really read on the host

    # ---- in-situ kernel times: the same step once more with an event pair around every kernel launch and the CUDA graphs
    # off (hdg_kernel_times; diagnostics, never part of `value`): where the device time of a step goes, kernel by kernel,
    # with the caches and the launch order of the real step.  One untimed step afterwards re-captures the graphs.
    insitu = None
    if args.insitu_steps > 0 and world == 1:
        eng.set_tuning("ktime", 1)
        eng.kernel_times()
        # (forcing as in the end-to-end arm just before, so that the time-extrapolated guesses see a continuous history)
        ims, _, its_insitu, step_no = timed_steps(args.insitu_steps, step_no, lambda kk: ts.step(kk, f_rhs, f_field=f_dev))
        kt = eng.kernel_times()
        eng.set_tuning("ktime", 0)
        ts.step(step_no, f_rhs)
        step_no += 1
        tot = sum(v[1] for v in kt.values())
        insitu = {"steps": args.insitu_steps, "iterations": {kk: v for kk, v in its_insitu.items() if kk != "per_step_tentative_pressure"},
                  "wall_ms_per_step_without_graphs": ims / args.insitu_steps,
                  "kernel_ms_per_step": tot / args.insitu_steps,
                  "graph_mode_ms_per_step": ms / args.steps,
                  "by_kernel": {name: {"launches_per_step": v[0] / args.insitu_steps, "ms_per_step": v[1] / args.insitu_steps,
                                       "us_per_launch": 1e3 * v[1] / max(v[0], 1), "share": v[1] / tot}
                                for name, v in sorted(kt.items(), key=lambda kv: -kv[1][1])[:30]}}

    # ---- cold start: the reference's behaviour (zero / Q^n initial guesses), a few steps of the same run ----------
    cold = None
    if args.cold_steps > 0:
        ts.warm_start = False
        eng.set_initial_guess(False)
        cms, _, its_cold, step_no = timed_steps(args.cold_steps, step_no)
        cold = {"value": work_units * args.cold_steps / (cms / 1e3), "unit": UNIT, "steps": args.cold_steps,
                "ms_per_step": cms / args.cold_steps, "iterations": its_cold}
    check_end = errors_and_checksums(step_no * dt)

    # ---- kernel probes: back-to-back launches between two CUDA events on the engine stream ------------------------
    def b2b(fn, nrep, nwarm=3):
        for _ in range(nwarm):
            fn()
        barrier()
        sa, sb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sa.record()
        for _ in range(nrep):
            fn()
        sb.record()
        torch.cuda.synchronize()
        return sa.elapsed_time(sb) / nrep

    xs, ys = eng.empty(2).normal_(), eng.empty(2)
    spmv_b2b_ms = b2b(lambda: eng.trace_spmv_dev(xs, ys), 20)
    Xf, Yf = eng.empty(0).normal_(), eng.empty(0)
    fimpl_b2b_ms = b2b(lambda: eng.fimpl_apply_dev(ts._Q_star.data, Xf, Yf, c0=1.0, c1=-dt), 10, 2)
    # forward elimination and back-substitution of the condensed mixed-Poisson solve (k_forward, k_back)
    Rp, lam = eng.empty(1).normal_(), eng.empty(2).normal_()
    Pf = eng.empty(1)
    fwd_ms = b2b(lambda: eng.forward_eliminate_dev(None, Rp, None, ys), 10, 2)
    back_ms = b2b(lambda: eng.back_substitute_dev(None, Rp, lam, Yf, Pf), 10, 2)
    del Xf, Yf, Rp, lam, Pf
    n_sweep, sweep_b2b_ms, sweep_err = 20, None, None
    try:
        sweep_b2b_ms = eng.tent_sweep_probe(dt, n_sweep)
    except Exception as exc:  # a measurement aid must not cost the bench line
        sweep_err = f"{type(exc).__name__}: {exc}"
    fp64_peak = eng.measure_fp64_peak()  # TFLOP/s, 8 independent DFMA chains per thread
    # condensation + assembly (K1-K3): re-run the set-up three times, event timers of k_condense / k_assemble
    t_before = eng.timers()
    for _ in range(3):
        eng.setup_poisson()
    t_after = eng.timers()
    condense_ms = (t_after["condense"][0] - t_before["condense"][0]) / 3
    assemble_ms = (t_after["assemble"][0] - t_before["assemble"][0]) / 3
    # ---- large time steps: the reference's default dt = 0.04 (src/driver.py:80-86), CFL = dt nx (41 at nx = 1024), and
    # a tenth of it; one cold-started step each with a bounded iteration budget (reported, never part of `value`)
    high_all = []
    if args.high_cfl_steps > 0:
        ts.warm_start = False
        eng.set_initial_guess(False)
        for hdt in args.high_cfl_dt:
            ts._dt = hdt
            ts.tentative_maxit = args.high_cfl_maxit
            ts.initialise(Q0, p0)
            stA = eng.tentative_stats()
            high = {"dt": hdt, "cfl": hdt * nx, "steps": args.high_cfl_steps, "unit": UNIT,
                    "tentative_maxit": args.high_cfl_maxit}
            try:
                hms, _, its_high, _ = timed_steps(args.high_cfl_steps, 0)
                high.update({"converged": True, "ms_per_step": hms / args.high_cfl_steps,
                             "value": work_units * args.high_cfl_steps / (hms / 1e3), "iterations": its_high,
                             "check": errors_and_checksums(args.high_cfl_steps * hdt)})
            except Exception as exc:  # a solve that does not converge must not cost the bench line
                high.update({"converged": False, "error": f"{type(exc).__name__}: {exc}"})
            stB = eng.tentative_stats()
            high["tentative_solver"] = {kk: stB[kk] - stA[kk] for kk in stB}
            high_all.append(high)
        ts._dt = dt
    high = high_all[0] if high_all else None

    comm_probe = None
    if world > 1:
        barrier()
        # one facet / cell halo exchange and one 2-slot all-reduce in isolation (back-to-back on the engine stream)
        fx, ar = eng.comm_probe(1, k + 2, nred=2, nrep=200)
        cx, _ = eng.comm_probe(0, (k + 2) * (k + 3), nred=0, nrep=200)
        comm_probe = {"facet_halo_us": max_over_ranks(fx), "cell_halo_us": max_over_ranks(cx),
                      "allreduce_2_slots_us": max_over_ranks(ar),
                      "what": "device time per call, 200 back-to-back calls between two events, max over ranks"}
        barrier()

    if rank == 0:
        peak, peak_src = load_peaks()
        b = k + 1
        nf_loc, nc_loc = eng.nf, eng.nc  # entities this rank's kernels run over (owned + ghost)
        nq1, npp, nl = (k + 2) * (k + 3) // 2, (k + 1) * (k + 2) // 2, 3 * (k + 1)
        na = 2 * nq1 + npp
        ms_step = ms / args.steps

        def hbm(name, launch_ms, nbytes, note=None, **extra):
            d = {"kernel": name, "bound": "hbm", "achieved": nbytes / launch_ms / 1e6, "peak": peak, "unit": "GB/s",
                 "frac": nbytes / launch_ms / 1e6 / peak, "peak_source": peak_src,
                 "algorithmic_bytes_per_launch": int(nbytes), "launch_ms": launch_ms}
            if note:
                d["note"] = note
            d.update(extra)
            return d

        kernels = {}
        nblocks = 5 * nf_loc
        spmv_bytes = nblocks * b * b * 8 + nblocks * 4 + 2 * b * nf_loc * 8
        kernels["k_cg_spmv"] = hbm(f"k_cg_spmv<{b}> (blocked-ELL trace SpMV of the CG; with N > 1 the launch includes this "
                                   "rank's facet-halo exchange)", spmv_b2b_ms, spmv_bytes, launches_timed=20,
                                   traffic=1353474384 if (world == 1 and nx == 1024 and k == 2) else None,
                                   traffic_source="ncu --set full, profiles/ncu_r1b_cg_spmv_raw.csv.gz")
        nm = k + 2
        f32 = bool(getattr(eng, "tuning", {}).get("tent_fp32", 1))
        # per facet: geometry 6 doubles + 5 ints (4 neighbour facets, orientation bits), rhs (double) and x, d (read), d,
        # xout (written) of NM entries each: FP32-stored iterate / correction (default) or FP64; the 4 neighbour facets' x
        # are L2 re-reads
        sweep_bytes = (6 * 8 + 5 * 4 + nm * 8 + 4 * nm * (4 if f32 else 8)) * nf_loc
        sweep_name = "k_tent_sweep32" if f32 else "k_tent_sweep"
        if sweep_b2b_ms:
            kernels[sweep_name] = hbm(f"{sweep_name}<{k}> (Chebyshev / facet-block-Jacobi sweep on the facet Schur complement "
                                      "of the tentative-velocity preconditioner)", sweep_b2b_ms, sweep_bytes,
                                      note="no switch over the local facet index since r2p (GG(e, e+j) = GG(0, j)); the r2l "
                                           "capture of the divergent version (profiles/r2/ncu_r2l_tent_sweep32_summary.txt) "
                                           "showed DRAM traffic = algorithmic bytes at 31 % occupancy",
                                      launches_timed=n_sweep, traffic=SWEEP_TRAFFIC.get((f32, world, nx, k)),
                                      traffic_source=SWEEP_TRAFFIC_SOURCE)
        else:
            kernels[sweep_name] = {"error": sweep_err}
        dfma_per_cell = {1: 7 * 48 + 3 * 220, 2: 16 * 80 + 3 * 350, 3: 36 * 120 + 3 * 735, 4: 64 * 168 + 3 * 1176}[k]
        tf = 2.0 * dfma_per_cell * nc_loc / fimpl_b2b_ms / 1e9
        kernels["k_fimpl"] = {"kernel": f"k_fimpl<{k},upwind> (matrix-free advection + flux + penalty operator: residuals, "
                                        "explicit terms; the Krylov iteration runs k_fimpl_c)",
                              "bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                              "peak_source": "self-measured FP64 FMA rate (hdg_measure_fp64_peak; MEASURED_PEAKS.json has no "
                                             "FP64 entry)", "dfma_per_cell": dfma_per_cell, "launch_ms": fimpl_b2b_ms,
                              "launches_timed": 10,
                              "hbm_frac": (4 * 2 * nq1 * 8) * nc_loc / fimpl_b2b_ms / 1e6 / peak}
        # operator of the Krylov iteration since r2p: k_fimpl_c (penalty-free, one thread per (cell, component), Q* from
        # the table of k_fimpl_pre).  Timed in situ (event pairs around every launch of a real step); per cell it moves
        # the table (2 NQ + 3 NQF doubles), x, z (read) and y (written) of 2 NQ1 doubles each, 6 geometry doubles and
        # 6 neighbour ints -- the neighbours' x are L2 re-reads.  FP64 work next to it: ~2 000 DFMA per cell at k = 2.
        nqv, nqf = {1: 7, 2: 16, 3: 36, 4: 64}[k], (3 * k + 5) // 2
        insitu_k = (insitu or {}).get("by_kernel", {})
        if "k_fimpl_c" in insitu_k:
            fc_ms = insitu_k["k_fimpl_c"]["us_per_launch"] / 1e3
            fc_bytes = ((2 * nqv + 3 * nqf + 3 * 2 * nq1 + 6) * 8 + 6 * 4) * nc_loc
            kernels["k_fimpl_c"] = hbm(f"k_fimpl_c<{k},upwind> (matrix-free advection + upwind flux operator of the "
                                       "tentative-velocity Krylov iteration, one thread per (cell, component))", fc_ms,
                                       fc_bytes, launches_timed=int(insitu_k["k_fimpl_c"]["launches_per_step"] * insitu["steps"]),
                                       timed="in situ: event pair around every launch of a real step (insitu_kernel_times)",
                                       traffic=(1611003000 + 322940416) if (world == 1 and nx == 1024 and k == 2) else None,
                                       traffic_source="ncu --set full of one launch inside a step, "
                                                      "profiles/r2/ncu_r2s_fimpl_c_raw.csv.gz",
                                       note="HBM is the tighter of the two bounds (0.30 ms of bytes against 0.25 ms of FP64 "
                                            "work at k = 2); ncu r2s: DRAM 42 %, FP64 pipe 44 %, issue slots 59 % busy (1.35 "
                                            "UMOV per DFMA for the table immediates), 16 warps per SM, power-capped; "
                                            "constant-bank tables, L1 prefetch and TMA staging measured slower (DESIGN.md 4)")
            if k == 2:
                kernels["k_fimpl_c"]["dfma_per_cell"] = 2000
                kernels["k_fimpl_c"]["fp64_frac"] = 2.0 * 2000 * nc_loc / fc_ms / 1e9 / fp64_peak
        for kname, entry in kernels.items():
            if kname in insitu_k and "launch_ms" in entry:
                entry["insitu_launch_ms"] = insitu_k[kname]["us_per_launch"] / 1e3
        # condensation metric of BASELINE.json ("condensation % of roofline"): DESIGN.md 4 / SURVEY.md 8d bytes per cell
        # k >= 3: the defaults are the shared-memory-factor kernels of csrc/hdg_poisson_s.cuh (k = 3: condense and back,
        # k = 4: all three)
        n_cond = "k_condense_b" if k >= 3 else "k_condense"
        n_fwd = "k_forward_s" if k >= 4 else "k_forward"
        n_back = "k_back_s" if k >= 3 else "k_back"
        kernels["k_condense"] = hbm(f"{n_cond}<{k}> (local operator + Schur complement S_K, closed form)", condense_ms,
                                    (6 + nl * nl) * 8 * nc_loc, launches_timed=3,
                                    generic_route_flops_per_cell={1: 6030, 2: 28097, 3: 92586, 4: 246582}[k])
        kernels["k_assemble"] = hbm(f"k_assemble<{k}> (deterministic gather of S_K into the blocked-ELL trace matrix)",
                                    assemble_ms, (2 * 3 * b * b + 6 * b * b) * 8 * nf_loc, launches_timed=3)
        kernels["k_forward"] = hbm(f"{n_fwd}<{k}> + k_trace_rhs (forward elimination, recompute variant)", fwd_ms,
                                   (6 + npp + nl) * 8 * nc_loc + (2 * nl // 3 + nl // 3) * 8 * nf_loc, launches_timed=10)
        kernels["k_back"] = hbm(f"{n_back}<{k}> (back-substitution, recompute variant)", back_ms,
                                (6 + npp + nl + na) * 8 * nc_loc, launches_timed=10)
        # share of the step: launches per step (counted by the engine) x back-to-back launch time
        counts = {"k_cg_spmv": per_step_launches.get("k_cg_spmv", 0.0),
                  sweep_name: per_step_launches.get("k_tent_sweep32", 0.0) + per_step_launches.get("k_tent_sweep", 0.0),
                  "k_fimpl": per_step_launches.get("k_fimpl", 0.0), "k_fimpl_c": per_step_launches.get("k_fimpl_c", 0.0),
                  "k_forward": per_step_launches.get(n_fwd, 0.0),
                  "k_back": per_step_launches.get(n_back, 0.0)}
        for name, cnt in counts.items():
            if "launch_ms" in kernels.get(name, {}):
                kernels[name]["launches_per_step"] = cnt
                kernels[name]["share_of_step"] = cnt * kernels[name]["launch_ms"] / ms_step
        # where the in-situ step timed a kernel, its share is the measured one: device time of its launches / device time
        # of all launches of that step (the probes run on warm buffers; k_tent_sweep32 and the FP64 residual sweep
        # k_tent_sweep are separate kernels there)
        if insitu_k:
            for name in counts:
                if name in insitu_k and name in kernels and "launch_ms" in kernels[name]:
                    kernels[name]["launches_per_step_counted"] = kernels[name].get("launches_per_step")
                    kernels[name]["launches_per_step"] = insitu_k[name]["launches_per_step"]
                    kernels[name]["share_of_step"] = insitu_k[name]["share"]
                    kernels[name]["share_source"] = "in situ (insitu_kernel_times)"
        dominant = max((n_ for n_ in counts if "share_of_step" in kernels.get(n_, {})),
                       key=lambda n_: kernels[n_]["share_of_step"])
        roofline = dict(kernels[dominant])
        roofline.setdefault("traffic", None)
        roofline["selected_as"] = "largest share of the timed step among the probed kernels"
        other = {n_: v for n_, v in kernels.items() if n_ != dominant}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_sample(args, mesh.nc)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "precision_note": ("tentative-velocity solve = FP64 iterative refinement whose inner BiCGStab iterations run "
                               "in FP32 (HDG_TUNING=tent_mixed=1); everything else FP64"
                               if getattr(eng, "tuning", {}).get("tent_mixed") else
                               "FP64 throughout; only inside the right preconditioner of the tentative-velocity BiCGStab "
                               "are the Chebyshev sweep iterates and the inverse cell blocks STORED in FP32 (flexible "
                               "solution update; accepted on the FP64 preconditioned residual <= rtol ||x||)"),
            "config": {**workload_config(args, world), "tuning": dict(getattr(eng, "tuning", {})),
                       "initial_guess": f"time-extrapolated (degree {args.warm_order}); cold-start figure under cold_start"},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(np.prod(sQ)) * 8,
                    "d2h_bytes_per_step": (int(np.prod(sQ)) + int(np.prod(sp_))) * 8, "steps": e2e_steps,
                    "device_ms_per_step": e2e_ms / e2e_steps, "iterations": {kk: v for kk, v in its_e2e.items() if kk != "per_step_tentative_pressure"},
                    "copies": "pinned host buffers, engine copy stream, overlapped with the solver (hdg_upload_begin / "
                              "hdg_download_begin)", "host_result_absmax": e2e_probe},
            "gpu_launches": int(launches),
            "gpu_launches_per_step_by_kernel": {kk: v for kk, v in sorted(per_step_launches.items(), key=lambda kv: -kv[1])[:24]},
            "cuda_graph_replays": int(eng.graph_replays),
            "warm_start_restarts": int(eng.guess_restarts),
            "roofline": roofline,
            "other_kernels": other,
            "cpu_baseline": cpu,
            "insitu_kernel_times": insitu,
            "cold_start": cold,
            "high_cfl": high,
            "high_cfl_more": high_all[1:],
            "check": {"after_timed_region": check_main, "at_end": check_end},
            "iterations": {"trace_cg_per_solve": its_main["trace_cg_per_solve"],
                           "tentative_bicgstab_per_solve": its_main["tentative_per_solve"],
                           "per_step_tentative_pressure": ts.iteration_history[:args.warmup + args.steps],
                           "tentative_solver": {kk: st1[kk] - st0[kk] for kk in st1},
                           "tentative_mixed_precision": {kk: mx1[kk] - mx0[kk] for kk in mx1}},
            "comm": {**eng.comm_stats(), "transport": ("nvlink-p2p" if getattr(eng, "p2p", False) else "nccl")
                     if world > 1 else "none", "p2p_timeouts": eng.p2p_status(), "probe": comm_probe},
            "breakdown_ms_per_step": {lab: timers[lab][0] / args.steps for lab in
                                      ("bdm_projection", "tentative_velocity_solve", "forward_elimination",
                                       "trace_solve", "back_substitution")},
            "setup": {"seconds": setup_s, "device_memory_gb": mem_used_gb},
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
    ap.add_argument("--cpu-nx", type=int, default=256,
                    help="mesh size of the bounded CPU sample (256: 1/16 of the cells of the nx = 1024 workload)")
    ap.add_argument("--cpu-steps", type=int, default=1, help="timed steps of the cpu_baseline leg of the own arm")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end arm (0 = the same as --steps)")
    ap.add_argument("--warm-order", type=int, default=3, help="degree of the time extrapolation of the initial guesses")
    ap.add_argument("--insitu-steps", type=int, default=1,
                    help="extra steps with an event pair around every kernel (in-situ kernel times; 0 = skip)")
    ap.add_argument("--cold-steps", type=int, default=2, help="extra steps with the reference's cold starts (0 = skip)")
    ap.add_argument("--high-cfl-steps", type=int, default=1,
                    help="extra steps at the reference's default dt (src/driver.py:80-86), reported under high_cfl")
    ap.add_argument("--high-cfl-dt", type=float, nargs="+", default=[0.04, 0.004])
    ap.add_argument("--high-cfl-maxit", type=int, default=600, help="iteration budget of one tentative solve there")
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
