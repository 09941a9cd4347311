"""CPU execution of the per-cell / per-facet kernels of the condensed mixed-Poisson path (csrc/hdg_poisson.cuh:
`k_condense`, `k_assemble`, `k_forward`, `k_back`), compiled with g++ through
tests/host_kernels (test infrastructure; the engine has no CPU path), against the oracle's dense-LU restatement of
the Slate static condensation (`hdg_imex.py:123-135`).  The same kernels run on the GPU through the C-ABI in
tests/test_engine_poisson_gpu.py; this file keeps them covered when no GPU is around."""
import ctypes
import os
import sys

import numpy as np
import pytest

from incompressibleeulerhdg_b200.mesh import PeriodicSquareMesh, RandomAffineCells, UnitSquareMesh
from oracle.hdg_oracle import HDGOracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_kernels"))
import build as host_build  # noqa: E402

DP, IP = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
TAU = 1.0


def dp(a):
    return None if a is None else a.ctypes.data_as(DP)


def ip(a):
    return a.ctypes.data_as(IP)


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    return host_build.build("poisson_host.cpp", str(tmp_path_factory.mktemp("host_kernels")))


class HostMesh:
    """the engine's SoA copies of the mesh arrays (what hdg_create builds on the device)"""

    def __init__(self, mesh):
        nc = mesh.nc
        t32 = lambda a: np.ascontiguousarray(np.asarray(a).T, dtype=np.int32)  # noqa: E731
        self.nc, self.nf = nc, mesh.nf
        self.xy = np.ascontiguousarray(np.asarray(mesh.cell_xy, dtype=np.float64).transpose(1, 2, 0).reshape(6, nc))
        self.cell_facet, self.cell_flip = t32(mesh.cell_facet), t32(mesh.cell_flip)
        self.facet_cell, self.facet_local = t32(mesh.facet_cell), t32(mesh.facet_local)


def condense(lib, hm, k, suffix=""):
    nl = 3 * (k + 1)
    SK = np.full((nl * nl, hm.nc), np.nan)
    fn = getattr(lib, "ph_condense" + suffix)
    assert fn(k, hm.nc, dp(hm.xy), ip(hm.cell_flip), ctypes.c_double(TAU), dp(SK)) == 0
    return SK


# suffix "_s": the variants of csrc/hdg_poisson_s.cuh (Cholesky factor in shared memory, facet-blocked condensation;
# hdg_set_tuning "poisson_lsmem"), instantiated for k >= 3
VARIANTS = [(1, ""), (2, ""), (3, ""), (4, ""), (3, "_s"), (4, "_s")]


@pytest.mark.parametrize("k,suffix", VARIANTS)
def test_condense_kernel_matches_the_oracle(lib, k, suffix):
    mesh = RandomAffineCells(37)
    hm, o = HostMesh(mesh), HDGOracle(mesh, k)
    nl = 3 * (k + 1)
    SK = condense(lib, hm, k, suffix).T.reshape(mesh.nc, nl, nl)
    assert rel(SK, o.condensed_local()) < 1e-11
    if suffix:  # same operation order per entry as the register kernel: agreement to the last bits, symmetric by construction
        ref = condense(lib, hm, k).T.reshape(mesh.nc, nl, nl)
        assert rel(SK, ref) < 1e-14
        assert np.array_equal(SK, SK.transpose(0, 2, 1))


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("mesh_fn", [lambda: UnitSquareMesh(4, perturb=0.2), lambda: PeriodicSquareMesh(3, L=2 * np.pi)])
def test_assemble_kernel_matches_the_oracle(lib, k, mesh_fn):
    mesh = mesh_fn()
    hm, o = HostMesh(mesh), HDGOracle(mesh, k)
    b, nf = k + 1, mesh.nf
    SK = condense(lib, hm, k)
    val, col, dinv = np.zeros((5 * b * b, nf)), np.zeros((5, nf), np.int32), np.zeros((b * b, nf))
    assert lib.ph_assemble(k, hm.nc, nf, dp(SK), ip(hm.cell_facet), ip(hm.facet_cell), ip(hm.facet_local), dp(val),
                           ip(col), dp(dinv)) == 0
    blocks, cols = val.T.reshape(nf, 5, b, b), col.T
    P = np.zeros((nf * b, nf * b))
    for f in range(nf):
        for j in range(5):
            c = cols[f, j]
            P[f * b:(f + 1) * b, c * b:(c + 1) * b] += blocks[f, j]
    S = o.assemble_trace_matrix().toarray()
    assert rel(-P, S) < 1e-11
    D = dinv.T.reshape(nf, b, b)
    assert np.abs(np.einsum("fij,fjk->fik", D, blocks[:, 0]) - np.eye(b)).max() < 1e-11  # facet-block-Jacobi


@pytest.mark.parametrize("k,suffix", VARIANTS)
def test_forward_and_back_kernels_match_the_oracle(lib, k, suffix):
    ph_forward, ph_back = getattr(lib, "ph_forward" + suffix), getattr(lib, "ph_back" + suffix)
    ph_back_update = getattr(lib, "ph_back_update" + suffix)
    mesh = UnitSquareMesh(3, perturb=0.2)
    hm, o = HostMesh(mesh), HDGOracle(mesh, k)
    nc, nf, nl1 = mesh.nc, mesh.nf, k + 1
    rng = np.random.default_rng(3)
    Ru, Rp = rng.standard_normal((nc, 2, o.nQ1)), rng.standard_normal((nc, o.np_))
    lam = rng.standard_normal((nf, nl1))
    Ru_s = np.ascontiguousarray(Ru.transpose(1, 2, 0).reshape(2 * o.nQ1, nc))
    Rp_s, lam_s = np.ascontiguousarray(Rp.T), np.ascontiguousarray(lam.T)
    A, Bk, Ck, Dk = o.local_system()
    Rloc = np.concatenate([Ru.reshape(nc, o.nQ), Rp], axis=1)
    # forward elimination: gK = C_K A_K^-1 (Ru, Rp), in the global facet orientation
    gK = np.zeros((3 * nl1, nc))
    assert ph_forward(k, nc, dp(hm.xy), ip(hm.cell_flip), ctypes.c_double(TAU), dp(Ru_s), dp(Rp_s), dp(gK)) == 0
    x0 = np.linalg.solve(A, Rloc[:, :, None])[:, :, 0]
    assert rel(gK.T, np.einsum("nla,na->nl", Ck, x0)) < 1e-11
    # back-substitution: (u, phi) = A_K^-1 ((Ru, Rp) - B_K lam_K)
    uo, po = np.zeros((2 * o.nQ1, nc)), np.zeros((o.np_, nc))
    assert ph_back(k, nc, nf, dp(hm.xy), ip(hm.cell_flip), ip(hm.cell_facet), ctypes.c_double(TAU), dp(Ru_s),
                       dp(Rp_s), dp(lam_s), dp(uo), dp(po)) == 0
    xl = np.linalg.solve(A, (Rloc - np.einsum("nal,nl->na", Bk, lam.ravel()[o.trace_dofs()]))[:, :, None])[:, :, 0]
    assert rel(uo.reshape(2, o.nQ1, nc).transpose(2, 0, 1), xl[:, :o.nQ].reshape(nc, 2, o.nQ1)) < 1e-11
    assert rel(po.T, xl[:, o.nQ:]) < 1e-11
    # the same fused with the update of the caller (k_back_update): Qacc <- cq Qacc + cb Qbase + cu u, pacc <- cp pacc + phi,
    # partial sum of detJ phi_0 over the owned cells (the pressure shift of hdg_imex.py:471-478 follows from it)
    Qb, Qa, pa = rng.standard_normal(uo.shape), rng.standard_normal(uo.shape), rng.standard_normal(po.shape)
    for cq, cb, cu, cp in ((0.0, 1.0, 0.37, 0.0), (1.0, 1.0, 0.05, 1.0), (0.0, 0.0, 1.0, 0.0)):
        Qacc, pacc, part = Qa.copy(), pa.copy(), np.zeros(1)
        nc_own = nc - 2
        assert ph_back_update(k, nc, nc_own, nf, dp(hm.xy), ip(hm.cell_flip), ip(hm.cell_facet), ctypes.c_double(TAU),
                              dp(Ru_s), dp(Rp_s), dp(lam_s), ctypes.c_double(cq), ctypes.c_double(cb),
                              ctypes.c_double(cu), ctypes.c_double(cp), dp(Qb), dp(Qacc), dp(pacc), dp(part)) == 0
        assert rel(Qacc, cq * Qa + cb * Qb + cu * uo) < 1e-12
        assert rel(pacc, cp * pa + po) < 1e-12
        assert abs(part[0] - np.sum(o.detJ[:nc_own] * po[0, :nc_own])) < 1e-12 * np.abs(po).max()
