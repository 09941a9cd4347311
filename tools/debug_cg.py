#!/usr/bin/env python
"""Long Chorin run with per-step Krylov diagnostics (why did a warm-started trace CG stall?)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from incompressibleeulerhdg_b200.engine import HDGError  # noqa: E402
from incompressibleeulerhdg_b200.mesh import UnitSquareMesh  # noqa: E402
from incompressibleeulerhdg_b200.model_problems import TaylorGreen  # noqa: E402
from incompressibleeulerhdg_b200.timesteppers import IncompressibleEulerHDGImplicit  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
graphs = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
mesh = UnitSquareMesh(nx, perturb=0.1)
ts = IncompressibleEulerHDGImplicit(mesh, 2, 0.32 / nx, krylov_rtol=1e-12)
ts.engine.set_graphs(graphs)
prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
Q0, p0 = prob.initial_condition()
ts.initialise(Q0, p0)
f = prob.f_rhs()
# fail fast: cap the trace solve
orig = ts.pressure_solve.__wrapped__ if hasattr(ts.pressure_solve, "__wrapped__") else None


def pressure_solve(Rp, u, phi, lmbda):
    return ts.engine.poisson_apply_dev(None, Rp.data, None, u.data, phi.data, lmbda.data, rtol=ts.krylov_rtol,
                                       maxit=400, shift=True)


ts.pressure_solve = pressure_solve
print(json.dumps({"mg_info": ts.engine.mg_info()}), flush=True)
for k in range(nsteps):
    try:
        ts.step(k, f)
        sc = ts.engine.debug_scalars()
        print(json.dumps({"step": k, "its": ts.iteration_history[-1], "cg_rz_over_ref": sc["cg"]["rz"] / sc["cg"]["ref"],
                          "cg_ref": sc["cg"]["ref"], "restarts": ts.engine.guess_restarts, "bicg_rr_over_bb": sc["bicgstab"]["rr"] / sc["bicgstab"]["bb"]}),
              flush=True)
    except HDGError as exc:
        sc = ts.engine.debug_scalars()
        nan = {n: bool(torch.isnan(t.data).any()) for n, t in
               (("Q", ts.Q), ("Qt", ts._Q_tentative), ("Rp", ts._Rp), ("lmbda", ts._lmbda), ("lmbda_prev", ts._lmbda_prev),
                ("u", ts._u), ("phi", ts._phi))}
        print(json.dumps({"step": k, "error": str(exc), "scalars": sc, "nan": nan,
                          "norm_Rp": float(ts._Rp.data.norm()), "norm_lmbda": float(ts._lmbda.data.norm())}), flush=True)
        break
