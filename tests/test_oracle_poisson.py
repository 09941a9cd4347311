"""CPU tests of the oracle itself (mixed-Poisson path): structural invariants and convergence.

The reference has no tests or golden vectors; these pin the oracle on the
invariants derived from the forms `hdg_imex.py:123-127,333-351` and on an analytic solution.
"""
import numpy as np
import pytest

from incompressibleeulerhdg_b200.mesh import PeriodicSquareMesh, UnitDiskMesh, UnitSquareMesh
from oracle.hdg_oracle import HDGOracle


@pytest.mark.parametrize("k", [1, 2, 3])
def test_local_invariants(k):
    m = UnitSquareMesh(3, perturb=0.2)
    o = HDGOracle(m, k)
    SK = o.condensed_local()
    assert np.abs(SK - SK.transpose(0, 2, 1)).max() < 1e-11
    ones = np.tile(np.eye(k + 1)[0], 3)
    assert np.abs(SK @ ones).max() < 1e-11  # S_K 1 = 0
    assert np.linalg.eigvalsh(SK).max() < 1e-11  # negative semi-definite
    M = o.local_blocks()["M"]
    assert np.abs(M - o.detJ[:, None, None] * np.eye(o.nQ)).max() < 1e-14  # orthonormal basis (H6)
    K, _ = o.assemble_monolithic()
    z = np.concatenate([np.zeros(m.nc * o.nQ), o.const_p().ravel(), o.null_vector_trace().ravel()])
    assert np.abs(K @ z).max() < 1e-12  # null vector (0,1,1), hdg_imex.py:480-489


@pytest.mark.parametrize("mesh_fn", [lambda: UnitSquareMesh(6, perturb=0.15), lambda: PeriodicSquareMesh(4, L=2 * np.pi),
                                      lambda: UnitDiskMesh(2)])
@pytest.mark.parametrize("k", [1, 2])
def test_condensed_equals_monolithic(mesh_fn, k):
    m = mesh_fn()
    o = HDGOracle(m, k)
    rng = np.random.default_rng(3)
    Ru = rng.standard_normal((m.nc, 2, o.nQ1))
    Rp = rng.standard_normal((m.nc, o.np_))
    Rl = rng.standard_normal((m.nf, k + 1))
    a = o.solve_monolithic(Ru, Rp, Rl)
    b = o.solve_condensed(Ru, Rp, Rl)
    for x, y in zip(a, b):
        assert np.abs(x - y).max() <= 1e-9 * max(1.0, np.abs(x).max())


def test_weak_divergence_is_consistent():
    m = UnitSquareMesh(5, perturb=0.1)
    o = HDGOracle(m, 2)
    Q = np.random.default_rng(0).standard_normal((m.nc, 2, o.nQ1))
    wd = o.weak_divergence(Q)
    z = np.zeros
    assert abs(o.consistency_defect(z((m.nc, 2, o.nQ1)), wd, z((m.nf, 3)))) < 1e-11
    # the by-parts form used for the pressure-reconstruction rhs (hdg_imex.py:204-207) is an identity
    wd2 = o.weak_divergence_fun(o.eval_Q(Q), o.eval_Q_facet(Q))
    assert np.abs(wd - wd2).max() < 1e-12
    # Chorin's broken divergence (hdg_implicit.py:145) is *not* consistent in general
    assert abs(o.consistency_defect(z((m.nc, 2, o.nQ1)), o.cell_divergence(Q), z((m.nf, 3)))) > 1e-3


@pytest.mark.parametrize("k", [1, 2])
def test_poisson_convergence(k):
    """-lap(phi) = f with homogeneous Neumann data: phi converges at rate k+1"""
    errs = []
    for nx in (4, 8):
        m = UnitSquareMesh(nx, perturb=0.1)
        o = HDGOracle(m, k)
        xp = o.phys_points()
        f = 2 * np.pi ** 2 * np.cos(np.pi * xp[..., 0]) * np.cos(np.pi * xp[..., 1])
        Rp = o.detJ[:, None] * np.einsum("q,nq,aq->na", o.wq, f, o.phiP)
        u, p, lam = o.solve_condensed(np.zeros((m.nc, 2, o.nQ1)), Rp, np.zeros((m.nf, k + 1)))
        errs.append(o.l2_error_p(p, lambda x, y: np.cos(np.pi * x) * np.cos(np.pi * y)))
    rate = np.log2(errs[0] / errs[1])
    assert rate > k + 1 - 0.35, (errs, rate)


def test_timestepper_convergence_rates():
    """observed rates of the IMEX SSP2(3,3,2) oracle against the exact Taylor-Green solution Psi(t) Q_s
    (`model_problems.py:56-105`), k = 1, nx = 4 -> 8: velocity -> k+2, pressure -> k+1.  The same run on the
    engine is `tests/test_driver.py::test_driver_observed_convergence_rates` (nx = 8 -> 16)."""
    from incompressibleeulerhdg_b200.mesh import UnitSquareMesh
    from oracle.timesteppers import IMEXOracle, TaylorGreenOracle

    prob = TaylorGreenOracle("exponential", 0.5)
    T, k, errs = 0.05, 1, []
    for nx in (4, 8):
        orc = IMEXOracle(UnitSquareMesh(nx), k, 0.0125, tableau="imex_ssp2_332")
        Q, p = orc.solve(prob, T)
        o, psi = orc.o, prob.psi(T)
        Qe = o.interpolate_cell(lambda x, y: tuple(psi * c for c in prob.Q_stationary(x, y)), "Q")
        pe = o.interpolate_cell(lambda x, y: psi ** 2 * prob.p_stationary(x, y), "p")
        pe = pe - o.integral_p(pe) * o.const_p()
        errs.append((o.l2_norm_Q(Q - Qe), float(np.sqrt(np.einsum("n,na->", o.detJ, (p - pe) ** 2)))))
    rate_Q = np.log2(errs[0][0] / errs[1][0])
    rate_p = np.log2(errs[0][1] / errs[1][1])
    assert abs(rate_Q - 2.651) < 0.02 and abs(rate_p - 1.905) < 0.02, (errs, rate_Q, rate_p)
