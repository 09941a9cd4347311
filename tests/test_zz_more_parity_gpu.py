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
