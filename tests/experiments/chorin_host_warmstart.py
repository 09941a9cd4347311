"""Development tool (not collected by pytest): BiCGStab iterations per Chorin step of the host-compiled device
kernels (tests/test_chorin_host.py) in the regime of bench.py -- dt = 0.32 / nx, tentative solve warm-started from
Q^n + (Q~^{n-1} - Q^{n-1}), rtol 1e-12 -- with and without the cell-block advection preconditioner.

    python tests/experiments/chorin_host_warmstart.py [nx=8] [k=2] [steps=6] [sweeps=8] [alpha=1]
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import test_chorin_host as C  # noqa: E402
from incompressibleeulerhdg_b200.mesh import UnitSquareMesh  # noqa: E402
from oracle.timesteppers import ChorinOracle, TaylorGreenOracle  # noqa: E402


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    sweeps = int(sys.argv[4]) if len(sys.argv) > 4 else 8
    alpha = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
    out = tempfile.mkdtemp()
    libs = {n: C.host_build.build(n + "_host.cpp", out) for n in ("poisson", "mg", "flow", "tent")}
    mesh, dt = UnitSquareMesh(nx, perturb=0.1), 0.32 / nx
    orc = ChorinOracle(mesh, k, dt)
    prob = TaylorGreenOracle("exponential", 0.5)
    for cb in (False, True):
        hc = C.HostChorin(libs, mesh, k, dt, "upwind")
        hc.tent = C.HostTentative(libs["tent"], mesh, k, alpha=alpha, sweeps=sweeps)
        Q = C.soa(orc.initial_state(prob)[0])
        dQt, its = np.zeros_like(Q), []
        for n in range(steps):
            f = C.soa(orc.interp_Q(prob.f_rhs(n * dt)))
            Qstar = hc.project_bdm(Q)
            Qt, it = hc.tent.solve(Qstar, dt, True, Q + dt * f, 1e-12, cb, x0=Q + dQt)
            its.append(it)
            Rp = np.zeros((hc.np_, mesh.nc))
            assert libs["flow"].fh_weak_div(k, mesh.nc, C.dp(hc.hm.xy), C.ip(hc.tent.nbr), C.ip(hc.tent.nbr_e), C.dp(Qt),
                                            C.cd(-1.0 / dt), 0, C.dp(Rp)) == 0
            u, phi, lam, _ = hc.poisson_apply(Rp)
            dQt = Qt - Q
            Q = Qt + dt * u
        print(f"nx={nx} k={k} sweeps={sweeps} alpha={alpha:g} cell blocks {int(cb)}: BiCGStab iterations per step {its}", flush=True)


if __name__ == "__main__":
    main()
