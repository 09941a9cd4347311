"""GPU parity of the per-cell Poisson kernels that keep the Cholesky factor in shared memory (csrc/hdg_poisson_s.cuh:
`k_condense_b`, `k_forward_s`, `k_back_s`, `k_back_update_s`; k >= 3, hdg_set_tuning "poisson_lsmem") against the oracle
and against the register kernels of csrc/hdg_poisson.cuh, through the C-ABI.  The same kernels are executed on the CPU
in tests/test_poisson_host.py.

Tolerance: BASELINE.json north_star asks for relative 1e-10 per solve on velocity, pressure, trace.
"""
import numpy as np
import pytest

from incompressibleeulerhdg_b200.engine import HDGEngine
from incompressibleeulerhdg_b200.mesh import RandomAffineCells, UnitSquareMesh
from oracle.hdg_oracle import HDGOracle
from conftest import require_degree

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("k", [3, 4])
def test_condensation_with_shared_factor_matches_oracle(k):
    """257 cells leave the last block of either launch (32 or 64 threads) partially filled"""
    require_degree(k)
    m = RandomAffineCells(257)
    ref = HDGOracle(m, k).condensed_local()
    eng = HDGEngine(m, k)
    out = {}
    for mask in (0, 1):
        eng.set_tuning("poisson_lsmem", mask)
        eng.setup_poisson(keep_local=True)
        out[mask] = eng.get_local_schur()
        assert rel(out[mask], ref) < 1e-11, mask
    assert rel(out[1], out[0]) < 1e-13
    assert np.array_equal(out[1], out[1].transpose(0, 2, 1))  # every facet pair computed once and mirrored


@pytest.mark.parametrize("k", [3, 4])
def test_poisson_apply_with_shared_factor_matches_oracle(k):
    require_degree(k)
    m = UnitSquareMesh(9, perturb=0.15)  # 162 cells: 5 blocks of 32 + 2 cells, 2 blocks of 64 + 34 cells
    o = HDGOracle(m, k)
    rng = np.random.default_rng(5)
    Ru = rng.standard_normal((m.nc, 2, o.nQ1))
    Rp = rng.standard_normal((m.nc, o.np_))
    Rl = rng.standard_normal((m.nf, k + 1))
    Rl[:, 0] -= o.consistency_defect(Ru, Rp, Rl) / m.nf
    Qo, po, lo = o.solve_condensed(Ru, Rp, Rl)
    Qa, Qb, pa = rng.standard_normal(Ru.shape), rng.standard_normal(Ru.shape), rng.standard_normal(Rp.shape)
    eng = HDGEngine(m, k)
    res = {}
    for mask in (0, 7):
        eng.set_tuning("poisson_lsmem", mask)
        eng.setup_poisson()
        before = eng.kernel_counts()
        Q, p, l, its = eng.poisson_apply_host(Ru, Rp, Rl, rtol=1e-13, maxit=20000)
        assert its > 0
        assert rel(Q, Qo) < 1e-10 and rel(p, po) < 1e-10 and rel(l, lo) < 1e-10, mask
        # back-substitution fused with the caller's update: Qacc <- Qacc + Qbase + 0.37 u, pacc <- pacc + phi
        dRu, dRp, dRl = eng.upload(0, Ru), eng.upload(1, Rp), eng.upload(2, Rl)
        Qacc, Qbase, pacc, dl = eng.upload(0, Qa), eng.upload(0, Qb), eng.upload(1, pa), eng.empty(2)
        eng.poisson_apply_update_dev(dRu, dRp, dRl, Qacc, pacc, dl, cq=1.0, cb=1.0, Q_base=Qbase, cu=0.37, cp=1.0,
                                     rtol=1e-13, maxit=20000)
        Qu, pu, lu = eng.download(0, Qacc), eng.download(1, pacc), eng.download(2, dl)
        assert rel(Qu, Qa + Qb + 0.37 * Qo) < 1e-10 and rel(pu, pa + po) < 1e-10 and rel(lu, lo) < 1e-10, mask
        after = eng.kernel_counts()
        launched = {n for n in after if after[n] > before.get(n, 0)}
        if mask:
            assert {"k_forward_s", "k_back_s", "k_back_update_s"} <= launched
            assert not {"k_forward", "k_back", "k_back_update"} & launched
        else:
            assert {"k_forward", "k_back", "k_back_update"} <= launched
        res[mask] = (Q, p, l, Qu, pu)
    for a, b in zip(res[7], res[0]):
        assert rel(a, b) < 1e-11
