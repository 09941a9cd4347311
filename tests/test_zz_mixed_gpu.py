"""Mixed-precision tentative-velocity solver (opt-in, hdg_set_tuning("tent_mixed", 1); run_tentative_mixed: FP64 iterative refinement around an FP32 BiCGStab):
same fields as the all-FP64 solver and as the oracle's sparse-direct solve, and the refinement really is the path taken."""

import numpy as np
import pytest

from conftest import require_degree
from incompressibleeulerhdg_b200 import timesteppers as TS
from incompressibleeulerhdg_b200.mesh import PeriodicSquareMesh, UnitSquareMesh
from incompressibleeulerhdg_b200.model_problems import TaylorGreen
from oracle.timesteppers import ChorinOracle, TaylorGreenOracle

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize("k,nx,flux", [(1, 8, "upwind"), (2, 8, "upwind"), (2, 6, "centered"), (3, 4, "upwind")])
def test_mixed_solver_matches_fp64_solver_and_oracle(k, nx, flux):
    require_degree(k)
    mesh, dt, nt = UnitSquareMesh(nx, perturb=0.1), 0.32 / nx, 3
    Qo, po = ChorinOracle(mesh, k, dt, flux=flux).solve(TaylorGreenOracle("exponential", 0.5), nt * dt)
    out = {}
    for mixed in (1, 0):
        ts = TS.IncompressibleEulerHDGImplicit(mesh, k, dt, flux=flux, krylov_rtol=1e-13)
        ts.engine.set_tuning("tent_mixed", mixed)
        prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
        Q, p = ts.solve(*prob.initial_condition(), None, prob.f_rhs(), nt * dt)
        out[mixed] = (Q.to_host(), p.to_host(), ts.niter_tentative.value, ts.engine.mixed_stats())
        assert rel(out[mixed][0], Qo) < 1e-10 and rel(out[mixed][1], po) < 1e-10, (mixed, k, flux)
    st = out[1][3]
    print(f"k={k} {flux}: iterations per solve mixed {out[1][2]:.1f} / fp64 {out[0][2]:.1f}; {st}")
    assert st["solves"] == nt and st["handed_to_fp64"] == 0 and st["inner_fp32_iterations"] > 0
    assert out[0][3]["solves"] == 0
    assert rel(out[1][0], out[0][0]) < 1e-10


def test_mixed_solver_hands_large_time_steps_to_the_fp64_solver():
    """CFL 4: the FP32 refinement may stagnate; the result must still be the oracle's"""
    k, nx = 2, 8
    require_degree(k)
    mesh, dt, nt = UnitSquareMesh(nx, perturb=0.1), 4.0 / nx, 2
    Qo, po = ChorinOracle(mesh, k, dt).solve(TaylorGreenOracle("exponential", 0.5), nt * dt)
    ts = TS.IncompressibleEulerHDGImplicit(mesh, k, dt, krylov_rtol=1e-13)
    ts.engine.set_tuning("tent_mixed", 1)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q, p = ts.solve(*prob.initial_condition(), None, prob.f_rhs(), nt * dt)
    print(ts.engine.mixed_stats(), ts.engine.tentative_stats())
    assert rel(Q.to_host(), Qo) < 1e-10 and rel(p.to_host(), po) < 1e-10


def test_mixed_solver_periodic_mesh_warm_start():
    k = 2
    require_degree(k)
    mesh, dt, nt = PeriodicSquareMesh(6, L=1.0), 0.05, 4
    Qo, po = ChorinOracle(mesh, k, dt).solve(TaylorGreenOracle("exponential", 0.5), nt * dt)
    ts = TS.IncompressibleEulerHDGImplicit(mesh, k, dt, krylov_rtol=1e-13, warm_start=True)
    ts.engine.set_tuning("tent_mixed", 1)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q, p = ts.solve(*prob.initial_condition(), None, prob.f_rhs(), nt * dt)
    assert rel(Q.to_host(), Qo) < 1e-10 and rel(p.to_host(), po) < 1e-10
    assert ts.engine.mixed_stats()["handed_to_fp64"] == 0
