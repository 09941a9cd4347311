"""GPU parity of the device timesteppers (Python mirrors of hdg_implicit.py / hdg_imex.py calling
the C-ABI) against the oracle's direct-solver restatement.  Tolerance 1e-10 relative per field
(BASELINE.json north_star); a few steps, small meshes."""
import numpy as np
import pytest

from conftest import require_degree
from incompressibleeulerhdg_b200.mesh import UnitSquareMesh
from incompressibleeulerhdg_b200.model_problems import TaylorGreen
from incompressibleeulerhdg_b200 import timesteppers as TS
from oracle.timesteppers import ChorinOracle, IMEXOracle, TaylorGreenOracle

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("k,nx", [(1, 6), (2, 5)])
@pytest.mark.parametrize("flux", ["upwind", "centered"])
def test_chorin_matches_oracle(k, nx, flux):
    require_degree(k)
    m = UnitSquareMesh(nx, perturb=0.1)
    dt, nt = 0.02, 3
    ts = TS.IncompressibleEulerHDGImplicit(m, k, dt, flux=flux, use_projection_method=True, krylov_rtol=1e-13)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q0, p0 = prob.initial_condition()
    Q, p = ts.solve(Q0, p0, None, prob.f_rhs(), nt * dt)
    orc = ChorinOracle(m, k, dt, flux=flux)
    Qo, po = orc.solve(TaylorGreenOracle("exponential", 0.5), nt * dt)
    assert rel(Q.to_host(), Qo) < 1e-10
    assert rel(p.to_host(), po) < 1e-10


@pytest.mark.parametrize("name,cls", [
    ("imex_implicit", "IncompressibleEulerHDGIMEXImplicit"),
    ("imex_ars2_232", "IncompressibleEulerHDGIMEXARS2_232"),
    ("imex_ars3_443", "IncompressibleEulerHDGIMEXARS3_443"),
    ("imex_ssp2_332", "IncompressibleEulerHDGIMEXSSP2_332"),
    ("imex_ssp3_433", "IncompressibleEulerHDGIMEXSSP3_433"),
])
def test_imex_projection_matches_oracle(name, cls):
    k, nx, dt, nt = 1, 5, 0.02, 2
    require_degree(k)
    m = UnitSquareMesh(nx, perturb=0.1)
    ts = getattr(TS, cls)(m, k, dt, flux="upwind", use_projection_method=True, n_richardson=2, krylov_rtol=1e-13)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q0, p0 = prob.initial_condition()
    Q, p = ts.solve(Q0, p0, None, prob.f_rhs(), nt * dt)
    orc = IMEXOracle(m, k, dt, tableau=name, n_richardson=2)
    Qo, po = orc.solve(TaylorGreenOracle("exponential", 0.5), nt * dt)
    assert rel(Q.to_host(), Qo) < 1e-10
    assert rel(p.to_host(), po) < 1e-10


def test_imex_k2_ssp2():
    k, nx, dt, nt = 2, 4, 0.02, 2
    require_degree(k)
    m = UnitSquareMesh(nx, perturb=0.1)
    ts = TS.IncompressibleEulerHDGIMEXSSP2_332(m, k, dt, krylov_rtol=1e-13)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "constant", 0.5)
    Q0, p0 = prob.initial_condition()
    Q, p = ts.solve(Q0, p0, None, prob.f_rhs(), nt * dt)
    orc = IMEXOracle(m, k, dt, tableau="imex_ssp2_332")
    Qo, po = orc.solve(TaylorGreenOracle("constant", 0.5), nt * dt)
    assert rel(Q.to_host(), Qo) < 1e-10
    assert rel(p.to_host(), po) < 1e-10
