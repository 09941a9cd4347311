"""The compiled CPU baseline (oracle/cpu_ref/hdg_cpu_ref.cpp, bench.py's cpu_baseline / --impl reference arm) against
the numpy oracle that is pinned to the reference's forms: every stage and whole Chorin steps, to 1e-10."""

import numpy as np
import pytest

from incompressibleeulerhdg_b200.mesh import UnitDiskMesh, UnitSquareMesh
from oracle.cpu_ref import ChorinCpuRef
from oracle.hdg_oracle import HDGOracle
from oracle.timesteppers import ChorinOracle, TaylorGreenOracle


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))


def fields(o, seed=0):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((o.mesh.nc, 2, o.nQ1)), rng.standard_normal((o.mesh.nc, 2, o.nQ1))


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("flux", ["upwind", "centered"])
def test_stages_match_oracle(k, flux):
    mesh = UnitSquareMesh(5, perturb=0.1)
    o = HDGOracle(mesh, k, flux=flux)
    c = ChorinCpuRef(mesh, k, 0.05, flux=flux, rtol=1e-13)
    Q, X = fields(o)
    Qs = o.project_bdm(Q)
    assert rel(c.project_bdm(Q), Qs) < 1e-11
    assert rel(c.f_impl_apply(X, Qs), o.f_impl_apply(X, Qs)) < 1e-11
    assert rel(c.local_schur(), o.condensed_local()) < 1e-10
    rng = np.random.default_rng(1)
    Rp = rng.standard_normal((mesh.nc, o.np_))
    u, p, lam, its = c.poisson_solve(None, Rp, None)
    uo, po, lo = o.solve_condensed(np.zeros_like(Q), Rp, np.zeros((mesh.nf, o.nl1)))
    assert its > 0 and rel(u, uo) < 1e-10 and rel(p, po) < 1e-10 and rel(lam, lo) < 1e-10
    Ru, Rl = rng.standard_normal(Q.shape), rng.standard_normal((mesh.nf, o.nl1))
    u, p, lam, _ = c.poisson_solve(Ru, Rp, Rl)
    uo, po, lo = o.solve_condensed(Ru, Rp, Rl)
    assert rel(u, uo) < 1e-10 and rel(p, po) < 1e-10 and rel(lam, lo) < 1e-10


@pytest.mark.parametrize("k,adt", [(1, 0.02), (2, 0.02), (2, 0.05), (3, 0.02)])
def test_tentative_solve_matches_sparse_direct(k, adt):
    import scipy.sparse.linalg as spla

    mesh = UnitSquareMesh(6, perturb=0.1)
    ts = ChorinOracle(mesh, k, adt)
    c = ChorinCpuRef(mesh, k, adt, rtol=1e-13)
    _, b = fields(ts.o, 3)
    # a smooth advecting velocity (block-Jacobi BiCGStab is a baseline solver for CFL <~ 1, not a robust one)
    Q = ts.o.interpolate_cell(TaylorGreenOracle.Q_stationary, "Q") + 0.05 * fields(ts.o, 4)[0]
    Qs = ts.o.project_bdm(Q)
    xo = spla.splu(ts.tentative_matrix(Qs, adt)).solve(b.ravel()).reshape(b.shape)
    x, its = c.tentative_solve(Qs, adt, b)
    assert its > 0 and rel(x, xo) < 1e-10


@pytest.mark.parametrize("mesh_name,k,nx", [("square", 1, 8), ("square", 2, 16), ("disk", 2, 0)])
def test_chorin_steps_match_oracle(mesh_name, k, nx):
    mesh = UnitSquareMesh(nx, perturb=0.1) if mesh_name == "square" else UnitDiskMesh(1)
    dt, nt = (0.32 / nx if nx else 0.02), 2
    prob = TaylorGreenOracle("exponential", 0.5)
    Qo, po = ChorinOracle(mesh, k, dt).solve(prob, nt * dt)
    c = ChorinCpuRef(mesh, k, dt, rtol=1e-13)
    Q, p = c.solve(prob, nt * dt)
    assert rel(Q, Qo) < 1e-10 and rel(p, po) < 1e-10
    t = c.timers()
    assert t["timestep"]["calls"] == nt and t["tentative_velocity_solve"]["seconds"] > 0
