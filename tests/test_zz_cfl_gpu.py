"""Large time steps.  The reference's default is ``--dt 0.04`` (`src/driver.py:80-86`), CFL 10 - 40 on the benchmark
meshes, and it solves the tentative-velocity system with a direct LU (`hdg_implicit.py:126-129`) or GMRES + ILU
(`hdg_imex.py:224-228`), which do not care.  The engine's BiCGStab does: beyond CFL ~ 1 it needs hundreds of
iterations or stalls, and the solve then continues with the restarted flexible GMRES (run_fgmres, csrc/hdg_engine.cu).
Parity with the oracle's sparse-direct solves at CFL 1, 4, 10 (dt up to 1.25 on the 8 x 8 mesh)."""
import numpy as np
import pytest

import incompressibleeulerhdg_b200.timesteppers as TS
from conftest import require_degree
from incompressibleeulerhdg_b200.mesh import UnitSquareMesh
from incompressibleeulerhdg_b200.model_problems import TaylorGreen
from oracle.timesteppers import ChorinOracle, TaylorGreenOracle

pytestmark = [pytest.mark.gpu]


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("k,nx,cfl", [(2, 8, 1.0), (2, 8, 4.0), (2, 8, 10.0), (1, 8, 10.0), (3, 4, 4.0)])
def test_chorin_matches_the_oracle_at_large_cfl(k, nx, cfl):
    require_degree(k)
    mesh, dt, nt = UnitSquareMesh(nx, perturb=0.1), cfl / nx, 2
    ts = TS.IncompressibleEulerHDGImplicit(mesh, k, dt, flux="upwind", krylov_rtol=1e-12)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q, p = ts.solve(*prob.initial_condition(), None, prob.f_rhs(), nt * dt)
    Qo, po = ChorinOracle(mesh, k, dt).solve(TaylorGreenOracle("exponential", 0.5), nt * dt)
    st = ts.engine.tentative_stats()
    print(f"k={k} nx={nx} CFL {cfl:g} (dt={dt:g}): velocity {rel(Q.to_host(), Qo):.2e} pressure {rel(p.to_host(), po):.2e}; "
          f"tentative solver {st}")
    assert rel(Q.to_host(), Qo) < 1e-10 and rel(p.to_host(), po) < 1e-10


def test_fgmres_alone_matches_bicgstab():
    """``tent_krylov=2`` runs every tentative solve through the flexible GMRES: same fields as the default path"""
    k, nx = 2, 8
    require_degree(k)
    mesh, dt, nt = UnitSquareMesh(nx, perturb=0.1), 0.32 / nx, 2
    out = {}
    for mode in (0, 2):
        ts = TS.IncompressibleEulerHDGImplicit(mesh, k, dt, krylov_rtol=1e-13)
        ts.engine.set_tuning("tent_krylov", mode)
        prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
        Q, p = ts.solve(*prob.initial_condition(), None, prob.f_rhs(), nt * dt)
        out[mode] = (Q.to_host(), p.to_host(), ts.engine.tentative_stats())
    Qo, po = ChorinOracle(mesh, k, dt).solve(TaylorGreenOracle("exponential", 0.5), nt * dt)
    print(out[0][2], out[2][2])
    assert out[2][2]["fgmres_iterations"] > 0 and out[2][2]["bicgstab_iterations"] == 0
    assert out[0][2]["fallbacks"] == 0
    for mode in (0, 2):
        assert rel(out[mode][0], Qo) < 1e-10 and rel(out[mode][1], po) < 1e-10
