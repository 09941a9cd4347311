#!/usr/bin/env python
"""Solver-parameter sweep for the Chorin step on one GPU: tentative-solver Chebyshev sweeps and the
multigrid smoother counts.  Prints one JSON line per configuration.

usage: python tools/probe_params.py --nx 512 --sweeps 3 4 6 8 --mg 1,1 1,2 2,2 1,3
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from incompressibleeulerhdg_b200.mesh import UnitSquareMesh  # noqa: E402
from incompressibleeulerhdg_b200.model_problems import TaylorGreen  # noqa: E402
from incompressibleeulerhdg_b200.timesteppers import IncompressibleEulerHDGImplicit  # noqa: E402


def run(ts, f, steps, first):
    eng = ts.engine
    eng.reset_timers()
    ts.niter_pressure.reset()
    ts.niter_tentative.reset()
    for s in range(steps):
        ts.step(first + s, f)
    torch.cuda.synchronize()
    tm = eng.timers()
    return {"its_p": ts.niter_pressure.value, "its_t": ts.niter_tentative.value,
            "tent_ms": tm["tentative_velocity_solve"][0] / steps, "trace_ms": tm["trace_solve"][0] / steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=512)
    ap.add_argument("--k", type=int, default=2)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--sweeps", type=int, nargs="*", default=[3, 4, 6, 8])
    ap.add_argument("--mg", nargs="*", default=["1,1", "1,2", "2,2", "1,3", "2,1"])
    ap.add_argument("--ratio", type=float, nargs="*", default=[10.0])
    ap.add_argument("--rtol", type=float, default=1e-12)
    ap.add_argument("--cold", action="store_true", help="no warm starts (zero / Q^n initial guesses)")
    ap.add_argument("--minblocks", type=int, nargs="*", default=[])
    args = ap.parse_args()
    nx = args.nx
    mesh = UnitSquareMesh(nx, perturb=0.1)
    ts = IncompressibleEulerHDGImplicit(mesh, args.k, 0.32 / nx, krylov_rtol=args.rtol, warm_start=not args.cold)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q0, p0 = prob.initial_condition()
    ts.initialise(Q0, p0)
    f = prob.f_rhs()
    ts.step(0, f)
    step = 1
    for sw in args.sweeps:
        ts.engine.set_tentative_solver(1, sw)
        try:
            r = run(ts, f, args.steps, step)
        except Exception as exc:  # too few sweeps: BiCGStab stalls
            r = {"error": repr(exc)[:200]}
            ts.initialise(Q0, p0)
        step += args.steps
        print(json.dumps({"nx": nx, "tent_sweeps": sw, **r}), flush=True)
    ts.engine.set_tentative_solver(1, 8)
    for mb in args.minblocks:
        ts.engine.set_tuning("sweep_minblocks", mb)
        r = run(ts, f, args.steps, step)
        step += args.steps
        print(json.dumps({"nx": nx, "sweep_minblocks": mb, **r}), flush=True)
    ts.engine.set_tuning("sweep_minblocks", 5)
    H = ts.engine.hierarchy
    for spec in args.mg:
        sf, sc = (int(v) for v in spec.split(","))
        for ratio in args.ratio:
            ts.engine.mg_setup(hierarchy=H, smooth_fine=sf, smooth_coarse=sc, cheb_ratio=ratio)
            try:
                r = run(ts, f, args.steps, step)
            except Exception as exc:
                r = {"error": repr(exc)[:200]}
                ts.initialise(Q0, p0)
            step += args.steps
            print(json.dumps({"nx": nx, "mg": [sf, sc], "ratio": ratio, **r}), flush=True)


if __name__ == "__main__":
    main()
