"""Further GPU parity cases at the higher orders of BASELINE.json configs[3,4] (k = 3) -- complete time loops
against the oracle, 1e-10 relative (north_star).  Kernel-level parity for k = 1..4 is in
test_engine_*_gpu.py / test_tracer_gpu.py; these runs string the kernels together at k = 3."""
import numpy as np
import pytest

import incompressibleeulerhdg_b200.timesteppers as TS
from conftest import require_degree
from incompressibleeulerhdg_b200.functions import Expression
from incompressibleeulerhdg_b200.mesh import UnitSquareMesh
from incompressibleeulerhdg_b200.model_problems import TaylorGreen
from oracle.timesteppers import ChorinOracle, IMEXOracle, TaylorGreenOracle

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def tracer0(x, y):
    return np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y)


@pytest.mark.parametrize("flux", ["upwind", "centered"])
def test_chorin_k3_with_tracer(flux):
    k, nx, dt, nt = 3, 3, 0.02, 2
    require_degree(k)
    mesh = UnitSquareMesh(nx, perturb=0.1)
    ts = TS.IncompressibleEulerHDGImplicit(mesh, k, dt, flux=flux, krylov_rtol=1e-13)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q, p = ts.solve(*prob.initial_condition(), Expression(tracer0, 0), prob.f_rhs(), nt * dt)
    orc = ChorinOracle(mesh, k, dt, flux=flux)
    Qo, po = orc.solve(TaylorGreenOracle("exponential", 0.5), nt * dt, q_initial=tracer0)
    errs = rel(Q.to_host(), Qo), rel(p.to_host(), po), rel(ts.q_tracer.to_host(), orc.q_tracer)
    print(f"k={k} {flux}: velocity {errs[0]:.2e} pressure {errs[1]:.2e} tracer {errs[2]:.2e}")
    assert max(errs) < TOL


def test_imex_ars3_k2():
    k, nx, dt, nt = 2, 4, 0.02, 1
    require_degree(k)
    mesh = UnitSquareMesh(nx, perturb=0.1)
    ts = TS.IncompressibleEulerHDGIMEXARS3_443(mesh, k, dt, use_projection_method=True, n_richardson=2, krylov_rtol=1e-13)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q, p = ts.solve(*prob.initial_condition(), Expression(tracer0, 0), prob.f_rhs(), nt * dt)
    orc = IMEXOracle(mesh, k, dt, tableau="imex_ars3_443", n_richardson=2)
    Qo, po = orc.solve(TaylorGreenOracle("exponential", 0.5), nt * dt, q_initial=tracer0)
    errs = rel(Q.to_host(), Qo), rel(p.to_host(), po), rel(ts.q_tracer.to_host(), orc.q_tracer)
    print(f"ARS3(4,4,3) k={k}: velocity {errs[0]:.2e} pressure {errs[1]:.2e} tracer {errs[2]:.2e}")
    assert max(errs) < TOL


@pytest.mark.gpu
def test_warm_started_chorin_needs_no_trace_restart():
    """time-extrapolated initial guesses (opt-in, `warm_start=True`): same fields as the cold-started run, fewer
    iterations, and the trace solve never has to fall back to a zero guess (`hdg_guess_restarts`)"""
    from incompressibleeulerhdg_b200 import timesteppers as TS
    from incompressibleeulerhdg_b200.mesh import UnitSquareMesh
    from incompressibleeulerhdg_b200.model_problems import TaylorGreen

    k, nx, nt = 2, 12, 6
    require_degree(k)
    mesh, dt = UnitSquareMesh(nx, perturb=0.1), 0.32 / nx
    out = {}
    for warm in (False, True):
        ts = TS.IncompressibleEulerHDGImplicit(mesh, k, dt, krylov_rtol=1e-13, warm_start=warm)
        prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
        Q, p = ts.solve(*prob.initial_condition(), None, prob.f_rhs(), nt * dt)
        out[warm] = (Q.to_host(), p.to_host(), ts.niter_tentative.value, ts.niter_pressure.value, ts.engine.guess_restarts)
    assert out[True][4] == 0
    assert np.abs(out[True][0] - out[False][0]).max() < 1e-10 * np.abs(out[False][0]).max()
    assert np.abs(out[True][1] - out[False][1]).max() < 1e-10 * np.abs(out[False][1]).max()
    assert out[True][2] < out[False][2] and out[True][3] <= out[False][3]
