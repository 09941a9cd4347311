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


def test_imex_k3_ssp2_config3_element():
    """the element of BASELINE.json configs[3] (k = 3, IMEX SSP2(3,3,2) with the projection-preconditioned
    Richardson iteration, n_richardson = 2) on a mesh the oracle finishes in seconds"""
    k, nx, dt, nt = 3, 3, 0.02, 1
    require_degree(k)
    m = UnitSquareMesh(nx, perturb=0.1)
    ts = TS.IncompressibleEulerHDGIMEXSSP2_332(m, k, dt, use_projection_method=True, n_richardson=2, krylov_rtol=1e-13)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q0, p0 = prob.initial_condition()
    Q, p = ts.solve(Q0, p0, None, prob.f_rhs(), nt * dt)
    orc = IMEXOracle(m, k, dt, tableau="imex_ssp2_332", n_richardson=2)
    Qo, po = orc.solve(TaylorGreenOracle("exponential", 0.5), nt * dt)
    assert rel(Q.to_host(), Qo) < 1e-10
    assert rel(p.to_host(), po) < 1e-10


# ---- fully implicit (unsplit) stage: FGMRES on the monolithic system, hdg_implicit.py:153-186 ----------
def test_gamma_rows_are_consistent_with_the_poisson_solve():
    """Gamma(Q, p, l) applied to the solution of the condensed mixed-Poisson solve returns the
    constraint right-hand sides the solve was given (hdg_imex.py:342-351 vs :123-127)"""
    from incompressibleeulerhdg_b200.engine import HDGEngine
    from oracle.hdg_oracle import HDGOracle

    k = 2
    require_degree(k)
    m = UnitSquareMesh(6, perturb=0.15)
    o = HDGOracle(m, k)
    eng = HDGEngine(m, k)
    eng.setup_poisson()
    rng = np.random.default_rng(3)
    Ru = rng.standard_normal((m.nc, 2, o.nQ1))
    Rp = rng.standard_normal((m.nc, o.np_))
    Rl = rng.standard_normal((m.nf, k + 1))
    Rl[:, 0] -= o.consistency_defect(Ru, Rp, Rl) / m.nf  # consistent data: nothing is projected away
    dRu, dRp, dRl = eng.upload(0, Ru), eng.upload(1, Rp), eng.upload(2, Rl)
    Q, p, l = eng.empty(0), eng.empty(1), eng.empty(2)
    eng.poisson_apply_dev(dRu, dRp, dRl, Q, p, l, rtol=1e-14, shift=False)
    gp, gl = eng.empty(1), eng.empty(2)
    eng.gamma_apply_dev(Q, p, l, gp, gl)
    assert rel(eng.download(1, gp), Rp) < 1e-9
    assert rel(eng.download(2, gl), Rl) < 1e-9


@pytest.mark.parametrize("k,nx", [(1, 6), (2, 4)])
def test_fully_implicit_matches_oracle(k, nx):
    require_degree(k)
    m = UnitSquareMesh(nx, perturb=0.1)
    dt, nt = 0.02, 2
    ts = TS.IncompressibleEulerHDGImplicit(m, k, dt, flux="upwind", use_projection_method=False, krylov_rtol=1e-13)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q0, p0 = prob.initial_condition()
    Q, p = ts.solve(Q0, p0, None, prob.f_rhs(), nt * dt)
    orc = ChorinOracle(m, k, dt, flux="upwind", use_projection_method=False)
    Qo, po = orc.solve(TaylorGreenOracle("exponential", 0.5), nt * dt)
    assert ts._monolithic.last_iterations > 0
    assert rel(Q.to_host(), Qo) < 1e-10
    assert rel(p.to_host(), po) < 1e-9  # the pressure enters the velocity row scaled by dt


def test_imex_unsplit_matches_oracle():
    k, nx, dt = 1, 5, 0.02
    require_degree(k)
    m = UnitSquareMesh(nx, perturb=0.1)
    ts = TS.IncompressibleEulerHDGIMEXSSP2_332(m, k, dt, use_projection_method=False, krylov_rtol=1e-13)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q0, p0 = prob.initial_condition()
    Q, p = ts.solve(Q0, p0, None, prob.f_rhs(), dt)
    orc = IMEXOracle(m, k, dt, tableau="imex_ssp2_332", use_projection_method=False)
    Qo, po = orc.solve(TaylorGreenOracle("exponential", 0.5), dt)
    assert rel(Q.to_host(), Qo) < 1e-10
    assert rel(p.to_host(), po) < 1e-9


def test_cuda_graph_replay_is_bitwise_identical():
    """the Krylov iteration bodies replayed as CUDA graphs (default) give bit-for-bit the fields and
    iteration counts of the kernel-by-kernel launches (hdg_set_graphs)"""
    k = 2
    require_degree(k)
    m = UnitSquareMesh(10, perturb=0.1)
    dt, nt = 0.02, 3
    out = []
    for graphs in (True, False):
        ts = TS.IncompressibleEulerHDGImplicit(m, k, dt, krylov_rtol=1e-12)
        ts.engine.set_graphs(graphs)
        prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
        Q0, p0 = prob.initial_condition()
        Q, p = ts.solve(Q0, p0, None, prob.f_rhs(), nt * dt)
        out.append((Q.to_host(), p.to_host(), list(ts.iteration_history), ts.engine.graph_replays))
    assert out[0][3] > 0 and out[1][3] == 0
    assert out[0][2] == out[1][2]
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
