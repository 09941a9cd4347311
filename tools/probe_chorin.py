#!/usr/bin/env python
"""Chorin steps at several mesh sizes: iteration counts and per-label CUDA-event timers.

usage: python tools/probe_chorin.py --nx 128 256 512 [--k 2] [--steps 2] [--pc gtmg|jacobi]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from incompressibleeulerhdg_b200.mesh import UnitSquareMesh  # noqa: E402
from incompressibleeulerhdg_b200.model_problems import TaylorGreen  # noqa: E402
from incompressibleeulerhdg_b200.timesteppers import IncompressibleEulerHDGImplicit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, nargs="+", default=[128, 256])
    ap.add_argument("--k", type=int, default=2)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--pc", default="gtmg")
    ap.add_argument("--rtol", type=float, default=1e-12)
    ap.add_argument("--cfl", type=float, default=0.32)
    args = ap.parse_args()
    for nx in args.nx:
        t0 = time.time()
        mesh = UnitSquareMesh(nx, perturb=0.1)
        t_mesh = time.time() - t0
        dt = args.cfl / nx
        t0 = time.time()
        ts = IncompressibleEulerHDGImplicit(mesh, args.k, dt, flux="upwind", use_projection_method=True,
                                            krylov_rtol=args.rtol, preconditioner=args.pc)
        torch.cuda.synchronize()
        t_setup = time.time() - t0
        prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
        Q0, p0 = prob.initial_condition()
        ts.initialise(Q0, p0)
        f = prob.f_rhs()
        res = {"nx": nx, "k": args.k, "pc": ts.preconditioner, "mesh_s": t_mesh, "setup_s": t_setup}
        try:
            ts.step(0, f)
            ts.engine.reset_timers()
            ts.niter_pressure.reset()
            ts.niter_tentative.reset()
            torch.cuda.synchronize()
            t0 = time.time()
            for s in range(args.steps):
                ts.step(1 + s, f)
            torch.cuda.synchronize()
            res["s_per_step"] = (time.time() - t0) / args.steps
            res["its_pressure"] = ts.niter_pressure.value
            res["its_tentative"] = ts.niter_tentative.value
            res["timers_ms_per_step"] = {k: round(v[0] / args.steps, 3) for k, v in ts.engine.timers().items()}
            res["mem_GB"] = torch.cuda.max_memory_allocated() / 1e9
        except Exception as exc:  # report and go on with the next size
            res["error"] = repr(exc)
        print(json.dumps(res), flush=True)
        del ts
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
