"""GPU parity of the condensed mixed-Poisson path against the oracle (through the C-ABI).

Tolerance: BASELINE.json north_star asks for relative 1e-10 per solve on velocity, pressure, trace.
"""
import numpy as np
import pytest

from incompressibleeulerhdg_b200.engine import HDGEngine
from incompressibleeulerhdg_b200.mesh import PeriodicSquareMesh, RandomAffineCells, UnitDiskMesh, UnitSquareMesh
from oracle.hdg_oracle import HDGOracle
from conftest import require_degree

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def ell_to_dense(val, col, nf, b):
    S = np.zeros((nf * b, nf * b))
    for f in range(nf):
        for j in range(5):
            c = col[f, j]
            S[f * b:(f + 1) * b, c * b:(c + 1) * b] += val[f, j]
    return S


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_local_schur_matches_oracle(k):
    require_degree(k)
    m = RandomAffineCells(257)
    o = HDGOracle(m, k)
    eng = HDGEngine(m, k)
    eng.setup_poisson(keep_local=True)
    SK = eng.get_local_schur()
    ref = o.condensed_local()
    assert rel(SK, ref) < 1e-11
    # every default kernel computes an entry and its mirror image once: k <= 2 on and above the diagonal (k_condense),
    # k >= 3 facet block by facet block (k_condense_b, csrc/hdg_poisson_s.cuh)
    assert np.array_equal(SK, SK.transpose(0, 2, 1))


@pytest.mark.parametrize("k", [3, 4])
def test_condensation_kernel_variants_agree(k):
    """K >= 3 ships three condensation kernels: facet-blocked with the Cholesky factor in shared memory (k_condense_b,
    the default, tests/test_zz_lsmem_gpu.py), fully unrolled thread-per-cell with the factor in registers (k_condense,
    "poisson_lsmem" = 0) and a row loop with warp-uniform table loads ("condense_rows" = 1); the latter two are compared
    here and must both reproduce the oracle.  257 cells leave the last block of either launch partially filled."""
    require_degree(k)
    m = RandomAffineCells(257)
    ref = HDGOracle(m, k).condensed_local()
    eng = HDGEngine(m, k)
    eng.set_tuning("poisson_lsmem", 0)
    out = {}
    for variant in (0, 1):
        eng.set_tuning("condense_rows", variant)
        eng.setup_poisson(keep_local=True)
        out[variant] = eng.get_local_schur()
        assert rel(out[variant], ref) < 1e-11, variant
    assert rel(out[1], out[0]) < 1e-12


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("mesh_fn", [lambda: UnitSquareMesh(5, perturb=0.2), lambda: PeriodicSquareMesh(4, L=2 * np.pi)])
def test_trace_matrix_matches_oracle(k, mesh_fn):
    require_degree(k)
    m = mesh_fn()
    o = HDGOracle(m, k)
    eng = HDGEngine(m, k)
    eng.setup_poisson()
    val, col = eng.get_trace_matrix()
    P = ell_to_dense(val, col, m.nf, k + 1)
    S = o.assemble_trace_matrix().toarray()
    assert rel(-P, S) < 1e-11


@pytest.mark.parametrize("k", [1, 2, 3, 4])
@pytest.mark.parametrize("mesh_fn", [lambda: UnitSquareMesh(8, perturb=0.15), lambda: PeriodicSquareMesh(5, L=2 * np.pi),
                                      lambda: UnitDiskMesh(2)])
def test_poisson_apply_matches_oracle(k, mesh_fn):
    require_degree(k)
    m = mesh_fn()
    o = HDGOracle(m, k)
    eng = HDGEngine(m, k)
    eng.setup_poisson()
    rng = np.random.default_rng(11)
    Ru = rng.standard_normal((m.nc, 2, o.nQ1))
    Rp = rng.standard_normal((m.nc, o.np_))
    Rl = rng.standard_normal((m.nf, k + 1))
    # consistent data: remove the defect along the left null vector (the engine projects anyway)
    Rl[:, 0] -= o.consistency_defect(Ru, Rp, Rl) / m.nf
    Q, p, l, its = eng.poisson_apply_host(Ru, Rp, Rl, rtol=1e-13, maxit=20000)
    Qo, po, lo = o.solve_condensed(Ru, Rp, Rl)
    assert its > 0
    assert rel(Q, Qo) < RTOL and rel(p, po) < RTOL and rel(l, lo) < RTOL


def test_inconsistent_rhs_is_projected():
    """Chorin's rhs (hdg_implicit.py:145) is not in range(S); engine and oracle use the same projection"""
    k = 2
    m = UnitSquareMesh(6, perturb=0.1)
    o = HDGOracle(m, k)
    eng = HDGEngine(m, k)
    eng.setup_poisson()
    Qt = np.random.default_rng(5).standard_normal((m.nc, 2, o.nQ1))
    Rp = -10.0 * o.cell_divergence(Qt)
    Q, p, l, its = eng.poisson_apply_host(None, Rp, None, rtol=1e-13)
    Qo, po, lo = o.solve_condensed(np.zeros_like(Qt), Rp, np.zeros((m.nf, k + 1)))
    assert rel(Q, Qo) < RTOL and rel(p, po) < RTOL and rel(l, lo) < RTOL


def test_poisson_properties_large():
    """size-independent properties at a size the oracle cannot reach: linearity and residual"""
    import torch

    k = 2
    m = UnitSquareMesh(96, perturb=0.1)
    eng = HDGEngine(m, k)
    eng.setup_poisson()
    rng = np.random.default_rng(2)
    Rp1 = rng.standard_normal((m.nc, eng.np_))
    Rp2 = rng.standard_normal((m.nc, eng.np_))
    for R in (Rp1, Rp2):
        R[:, 0] -= (R[:, 0].sum()) / m.nc  # consistent: sum of mode-0 coefficients x const = 0 is enough here
    a = eng.poisson_apply_host(None, Rp1, None, rtol=1e-13, maxit=50000)
    b = eng.poisson_apply_host(None, Rp2, None, rtol=1e-13, maxit=50000)
    c = eng.poisson_apply_host(None, 2.0 * Rp1 - 3.0 * Rp2, None, rtol=1e-13, maxit=50000)
    for i in range(3):
        assert rel(c[i], 2.0 * a[i] - 3.0 * b[i]) < 1e-8
    # deterministic: bitwise identical on repetition
    a2 = eng.poisson_apply_host(None, Rp1, None, rtol=1e-13, maxit=50000)
    for i in range(3):
        assert np.array_equal(a[i], a2[i])
    assert a[3] == a2[3]
    # symmetry of the assembled operator through the SpMV: <x, P y> == <y, P x>
    x = torch.randn(eng.nl1 * m.nf, dtype=torch.float64, device="cuda")
    y = torch.randn_like(x)
    Px, Py = torch.empty_like(x), torch.empty_like(x)
    eng.trace_spmv_dev(x, Px)
    eng.trace_spmv_dev(y, Py)
    eng.synchronize()
    s1, s2 = float(torch.dot(y, Px)), float(torch.dot(x, Py))
    assert abs(s1 - s2) < 1e-10 * abs(s1)
    ones = torch.zeros_like(x)
    ones[: m.nf] = 1.0
    eng.trace_spmv_dev(ones, Px)
    eng.synchronize()
    assert float(Px.abs().max()) < 1e-9


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("mesh_fn", [lambda: UnitSquareMesh(16, perturb=0.15), lambda: PeriodicSquareMesh(12, L=2 * np.pi),
                                      lambda: UnitDiskMesh(3)])
def test_multigrid_pcg_matches_oracle(k, mesh_fn):
    """GTMG-preconditioned CG reaches the same solution in O(1) iterations"""
    require_degree(k)
    m = mesh_fn()
    o = HDGOracle(m, k)
    eng = HDGEngine(m, k)
    eng.setup_poisson()
    rng = np.random.default_rng(12)
    Ru = rng.standard_normal((m.nc, 2, o.nQ1))
    Rp = rng.standard_normal((m.nc, o.np_))
    Rl = rng.standard_normal((m.nf, k + 1))
    _, _, _, its_jac = eng.poisson_apply_host(Ru, Rp, Rl, rtol=1e-13, maxit=20000)
    H = eng.mg_setup()
    nl, lmax = eng.mg_info()
    assert nl == H.nlevels and 1.0 < lmax < 3.0
    Q, p, l, its = eng.poisson_apply_host(Ru, Rp, Rl, rtol=1e-13, maxit=200)
    Qo, po, lo = o.solve_condensed(Ru, Rp, Rl)
    assert its < 40 and its < its_jac
    assert rel(Q, Qo) < RTOL and rel(p, po) < RTOL and rel(l, lo) < RTOL


def test_p1_stiffness_is_galerkin_coarse_operator():
    """T^T (-S) T equals the P1 stiffness matrix (so rediscretisation == Galerkin, hdg_imex.py:101-106)"""
    from incompressibleeulerhdg_b200 import multigrid

    k = 2
    m = UnitSquareMesh(6, perturb=0.2)
    eng = HDGEngine(m, k)
    eng.setup_poisson()
    val, col = eng.get_trace_matrix()
    P = ell_to_dense(val, col, m.nf, k + 1)  # AoS numbering f*b+m
    b = k + 1
    perm = (np.arange(m.nf)[None, :] * b + np.arange(b)[:, None]).ravel()  # SoA index -> AoS index
    T = multigrid.trace_transfer(m, k).toarray()
    A = multigrid.p1_stiffness(m).toarray()
    G = T.T @ P[np.ix_(perm, perm)] @ T
    assert np.abs(G - A).max() < 1e-11 * np.abs(A).max()


@pytest.mark.parametrize("pc", ["jacobi", "gtmg"])
def test_initial_guess_does_not_change_the_result(pc):
    """hdg_set_initial_guess: starting the trace Krylov solve from a guess (here: the exact trace, a
    perturbed one and garbage) gives the zero-guess result to the solver tolerance; the tolerance is
    relative to the right-hand side, so a good guess only saves iterations"""
    import torch

    k = 2
    require_degree(k)
    m = UnitSquareMesh(12, perturb=0.15)
    eng = HDGEngine(m, k)
    eng.setup_poisson()
    if pc == "gtmg":
        eng.mg_setup()
    rng = np.random.default_rng(11)
    sQ, sp_, sl = eng.shapes()
    Rp = eng.upload(1, rng.standard_normal(sp_))
    Q, p, l = eng.empty(0), eng.empty(1), eng.empty(2)
    its0 = eng.poisson_apply_dev(None, Rp, None, Q, p, l, rtol=1e-13)
    Qz, pz, lz = Q.clone(), p.clone(), l.clone()
    eng.set_initial_guess(True)
    its_exact = eng.poisson_apply_dev(None, Rp, None, Q, p, l, rtol=1e-13)  # l holds the solution
    assert its_exact <= 1
    assert rel(p.cpu().numpy(), pz.cpu().numpy()) < 1e-11
    l.copy_(lz + 1e-6 * torch.randn_like(lz))
    its_near = eng.poisson_apply_dev(None, Rp, None, Q, p, l, rtol=1e-13)
    assert its_near < its0
    for a, b_ in ((Q, Qz), (p, pz), (l, lz)):
        assert rel(a.cpu().numpy(), b_.cpu().numpy()) < RTOL
    l.copy_(float(lz.abs().max()) * torch.randn_like(lz))  # a useless guess of the solution's own magnitude
    eng.poisson_apply_dev(None, Rp, None, Q, p, l, rtol=1e-13, maxit=100000)
    for a, b_ in ((Q, Qz), (p, pz), (l, lz)):
        assert rel(a.cpu().numpy(), b_.cpu().numpy()) < RTOL
    eng.set_initial_guess(False)
    assert eng.poisson_apply_dev(None, Rp, None, Q, p, l, rtol=1e-13) == its0
