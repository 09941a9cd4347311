"""Development tool (not collected by pytest): iteration count of the host port of the tentative solver
(tests/test_tent_host.py, real device kernels compiled with g++) against the number of Chebyshev sweeps, with and
without the cell-block advection preconditioner, and a cost model from the per-kernel times of
profiles/launches_r1k.md (us per launch at nx = 1024, k = 2: sweep 170, fimpl 721, xhat + moments + trhs 318,
BiCGStab vector kernels 600 per operator application; cell-block apply ~250 estimated from its 720 B/cell).

    python tests/experiments/tent_host_sweeps.py [nx=8] [k=2] [rtol=1e-6] [alpha=1]

The weight of the penalty relative to the mass matrix grows like a alpha / h^2 = cfl alpha / h: alpha = 1024 / nx on
the small mesh reproduces the stiffness of the nx = 1024 bench at the same advective CFL number.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import test_tent_host as T  # noqa: E402


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    rtol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-6
    alpha = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    lib = T.host_build.build("tent_host.cpp", tempfile.mkdtemp())
    mesh, o, Q0, Qs, adt = T._problem(k, nx, "upwind")
    b = Q0 + 0.01 * np.random.default_rng(11).standard_normal(Q0.shape)
    print(f"nx={nx} k={k} rtol={rtol:g} alpha={alpha:g}: sweeps, cell blocks, BiCGStab iterations, modelled ms per solve at nx=1024")
    for sweeps in (2, 3, 4, 6, 8, 10):
        for cb in (False, True):
            ht = T.HostTentative(lib, mesh, k, alpha=alpha, sweeps=sweeps)
            _, its = ht.solve(T.soa(Qs), adt, True, T.soa(b), rtol, cb)
            per_app = 170 * (sweeps + 1) + 721 + 318 + 600 + (250 if cb else 0)
            print(f"  sweeps {sweeps:2d}  cell blocks {int(cb)}  iterations {its:4d}  model {2 * its * per_app / 1e3:7.1f} ms",
                  flush=True)


if __name__ == "__main__":
    main()
