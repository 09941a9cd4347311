"""CPU execution of the kernels of the geometric-trace multigrid preconditioner (csrc/hdg_mg.cuh: the GPU apply of
`firedrake.GTMGPC`, `hdg_imex.py:138-169`), compiled with g++ through tests/host_kernels (test
infrastructure; the engine has no CPU path), on the blocked-ELL trace matrix produced by the host-compiled
`k_condense` + `k_assemble`.  A numpy port of `mg_apply` / `mg_vcycle` / `mg_smooth_csr` (csrc/hdg_engine.cu) strings
the kernels into the V-cycle; checked: every kernel against its numpy formula, symmetry of the V-cycle (what CG
needs), and that the multigrid-preconditioned CG reproduces the oracle's condensed solve in O(10) iterations where
facet-block-Jacobi CG needs many times more.  The CG vector kernels themselves live in hdg_engine.cu and run on
the GPU only (tests/test_engine_poisson_gpu.py)."""
import ctypes
import os
import sys

import numpy as np
import pytest

from incompressibleeulerhdg_b200 import multigrid
from incompressibleeulerhdg_b200.mesh import UnitSquareMesh
from oracle.hdg_oracle import HDGOracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_kernels"))
import build as host_build  # noqa: E402
from test_poisson_host import HostMesh, condense, dp, ip  # noqa: E402

cd = ctypes.c_double


@pytest.fixture(scope="module")
def libs(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host_kernels"))
    return host_build.build("poisson_host.cpp", out), host_build.build("mg_host.cpp", out)


def cheb_coefs(lmax, ratio, nsweeps):  # csrc/hdg_mg.cuh
    a, b = lmax / ratio, 1.1 * lmax
    theta, delta = 0.5 * (b + a), 0.5 * (b - a)
    sigma = theta / delta
    rho = 1.0 / sigma
    out = [(0.0, 1.0 / theta)]
    for _ in range(1, nsweeps):
        rho_new = 1.0 / (2.0 * sigma - rho)
        out.append((rho_new * rho, 2.0 * rho_new / delta))
        rho = rho_new
    return out


class Csr:
    def __init__(self, M):
        M = M.tocsr()
        M.sort_indices()
        self.n, self.m = M.shape
        self.rowptr = np.ascontiguousarray(M.indptr, np.int32)
        self.col = np.ascontiguousarray(M.indices, np.int32)
        self.val = np.ascontiguousarray(M.data, np.float64)


class HostGTMG:
    """numpy port of hdg_mg_setup / mg_apply / mg_vcycle / mg_smooth_csr on the host-compiled kernels"""

    def __init__(self, libs, mesh, k, ns_fine=1, ns_coarse=1, ratio=10.0):
        lp, self.lm = libs
        self.k, self.b, self.nf = k, k + 1, mesh.nf
        self.ns_fine, self.ns_coarse, self.ratio = ns_fine, ns_coarse, ratio
        hm = HostMesh(mesh)
        b, nf = self.b, self.nf
        SK = condense(lp, hm, k)
        self.val, self.col, self.dinv = np.zeros((5 * b * b, nf)), np.zeros((5, nf), np.int32), np.zeros((b * b, nf))
        assert lp.ph_assemble(k, hm.nc, nf, dp(SK), ip(hm.cell_facet), ip(hm.facet_cell), ip(hm.facet_local),
                              dp(self.val), ip(self.col), dp(self.dinv)) == 0
        H = multigrid.build_hierarchy(mesh, k)
        self.A, self.P = [Csr(a) for a in H.A], [Csr(p) for p in H.P]
        self.R = [Csr(p.T) for p in H.P]
        self.T, self.Tt = Csr(H.T), Csr(H.T.T)
        self.pinv, self.lmax = np.ascontiguousarray(H.pinv), list(H.lmax)
        self.dinv_l = []
        for a in self.A:
            d = np.zeros(a.n)
            assert self.lm.mh_csr_diag_inv(a.n, ip(a.rowptr), ip(a.col), dp(a.val), dp(d)) == 0
            self.dinv_l.append(d)
        # lambda_max(Dinv P) by power iteration (hdg_mg_setup does the same on the device)
        v = np.random.default_rng(0).standard_normal((b, nf))
        lam = 2.0
        for _ in range(100):
            v = self.blockjac(-self.residual(np.zeros((b, nf)), v))  # Dinv P v
            lam = np.linalg.norm(v)
            v /= lam
        self.fine_lmax = lam
        self.fd = np.zeros((b, nf))

    # -- kernels ---------------------------------------------------------------------------------------------
    def ell_cheb(self, bv, x, cdv, crv, zero):
        out = np.zeros((self.b, self.nf))
        assert self.lm.mh_ell_cheb(self.b, self.nf, dp(self.val), ip(self.col), dp(self.dinv), dp(bv), dp(x), dp(self.fd),
                                   dp(out), cd(cdv), cd(crv), int(zero)) == 0
        return out

    def residual(self, bv, x):
        r = np.zeros((self.b, self.nf))
        assert self.lm.mh_ell_residual(self.b, self.nf, dp(self.val), ip(self.col), dp(bv), dp(x), dp(r)) == 0
        return r

    def blockjac(self, r):
        z = np.zeros((self.b, self.nf))
        assert self.lm.mh_blockjac(self.b, self.nf, dp(self.dinv), dp(r), dp(z)) == 0
        return z

    def spmv(self, M, x, b=None, y=None, mode=0):
        y = np.zeros(M.n) if y is None else y
        assert self.lm.mh_csr_spmv(M.n, ip(M.rowptr), ip(M.col), dp(M.val), dp(np.ascontiguousarray(x.ravel())),
                                   dp(b), dp(y), mode) == 0
        return y

    # -- mg_smooth_csr / mg_vcycle / mg_apply --------------------------------------------------------------------
    def smooth_csr(self, l, b, x, zero):
        A, d = self.A[l], np.zeros(self.A[l].n)
        for j, (cdv, crv) in enumerate(cheb_coefs(self.lmax[l], self.ratio, self.ns_coarse)):
            out = np.zeros(A.n)
            assert self.lm.mh_csr_cheb(A.n, ip(A.rowptr), ip(A.col), dp(A.val), dp(self.dinv_l[l]), dp(b), dp(x), dp(d),
                                       dp(out), cd(cdv), cd(crv), int(zero and j == 0)) == 0
            x = out
        return x

    def vcycle(self, l, b):
        if l == len(self.A) - 1:
            x = np.zeros(self.A[l].n)
            assert self.lm.mh_dense_matvec(self.A[l].n, dp(self.pinv), dp(b), dp(x)) == 0
            return x
        x = self.smooth_csr(l, b, np.zeros(self.A[l].n), True)
        r = self.spmv(self.A[l], x, b=b, mode=2)
        xc = self.vcycle(l + 1, self.spmv(self.R[l], r))
        x = self.spmv(self.P[l], xc, y=x, mode=1)
        return self.smooth_csr(l, b, x, False)

    def apply(self, r):
        """z = M^-1 r: Chebyshev/block-Jacobi, P1 coarse correction, Chebyshev/block-Jacobi"""
        cc = cheb_coefs(self.fine_lmax, self.ratio, self.ns_fine)
        x = np.zeros((self.b, self.nf))
        for j, (cdv, crv) in enumerate(cc):
            x = self.ell_cheb(r, x, cdv, crv, j == 0)
        fr = self.residual(r, x)
        x0 = self.vcycle(0, self.spmv(self.Tt, fr))
        x = self.spmv(self.T, x0, y=np.ascontiguousarray(x.ravel()), mode=1).reshape(self.b, self.nf)
        for cdv, crv in cc:
            x = self.ell_cheb(r, x, cdv, crv, False)
        return x

    def matvec(self, x):  # P x = -(0 - P x)
        return -self.residual(np.zeros((self.b, self.nf)), x)

    def pcg(self, bvec, precond, rtol=1e-12, maxit=2000):
        """CG on P = -S with the constant mode projected out of the right-hand side (run_pcg_mg)"""
        bvec = bvec.copy()
        bvec[0] -= bvec[0].mean()
        x = np.zeros_like(bvec)
        r = bvec.copy()
        z = precond(r)
        p = z.copy()
        rz = rz0 = float(np.sum(r * z))
        its = 0
        while rz > rtol * rtol * rz0 and its < maxit:
            q = self.matvec(p)
            al = rz / float(np.sum(p * q))
            x += al * p
            r -= al * q
            z = precond(r)
            rz_new = float(np.sum(r * z))
            p = z + (rz_new / rz) * p
            rz = rz_new
            its += 1
        return x, its


@pytest.mark.parametrize("k", [1, 2, 3])
def test_fine_level_kernels_match_their_formulas(libs, k):
    mesh = UnitSquareMesh(4, perturb=0.15)
    mg = HostGTMG(libs, mesh, k)
    b, nf = mg.b, mg.nf
    o = HDGOracle(mesh, k)
    P = -o.assemble_trace_matrix().toarray()                      # oracle, dof = facet * b + mode
    perm = (np.arange(nf)[None, :] * b + np.arange(b)[:, None]).ravel()   # SoA [mode][facet] -> oracle numbering
    rng = np.random.default_rng(2)
    x, bv = rng.standard_normal((b, nf)), rng.standard_normal((b, nf))
    Px = P[np.ix_(perm, perm)] @ x.ravel()
    r_ref = bv.ravel() - Px
    assert np.abs(mg.residual(bv, x).ravel() - r_ref).max() < 1e-11 * np.abs(r_ref).max()
    D = np.linalg.inv(np.stack([P[f * b:(f + 1) * b, f * b:(f + 1) * b] for f in range(nf)]))
    z_ref = np.einsum("fij,jf->if", D, r_ref.reshape(b, nf))
    assert np.abs(mg.blockjac(r_ref.reshape(b, nf)) - z_ref).max() < 1e-10 * np.abs(z_ref).max()
    mg.fd[:] = rng.standard_normal((b, nf))
    d0 = mg.fd.copy()
    xout = mg.ell_cheb(bv, x, 0.3, 0.7, False)                    # d = cd d + cr Dinv r ; xout = x + d
    assert np.abs(mg.fd - (0.3 * d0 + 0.7 * z_ref)).max() < 1e-10 * np.abs(z_ref).max()
    assert np.abs(xout - (x + mg.fd)).max() < 1e-13


@pytest.mark.parametrize("k,nx", [(1, 8), (2, 8), (3, 4)])
def test_gtmg_preconditioned_cg_on_the_host(libs, k, nx):
    mesh = UnitSquareMesh(nx, perturb=0.1)
    mg = HostGTMG(libs, mesh, k)
    o = HDGOracle(mesh, k)
    b, nf = mg.b, mg.nf
    rng = np.random.default_rng(5)
    u, v = rng.standard_normal((b, nf)), rng.standard_normal((b, nf))
    Mu, Mv = mg.apply(u), mg.apply(v)
    assert abs(np.sum(v * Mu) - np.sum(u * Mv)) < 1e-10 * abs(np.sum(v * Mu))   # symmetric V-cycle
    # condensed mixed-Poisson solve with a pressure right-hand side (the Chorin / IMEX stage solve)
    Rp = rng.standard_normal((mesh.nc, o.np_))
    zero_u, zero_l = np.zeros((mesh.nc, 2, o.nQ1)), np.zeros((nf, b))
    _, _, _, parts = o.solve_condensed(zero_u, Rp, zero_l, return_parts=True)
    rhs = -np.ascontiguousarray(parts["r"].T)                     # P lam = -r  (P = -S), SoA [mode][facet]
    lam_mg, its_mg = mg.pcg(rhs, mg.apply)
    lam_bj, its_bj = mg.pcg(rhs, mg.blockjac)
    S = parts["S"]
    for lam in (lam_mg, lam_bj):
        res = S @ lam.T.ravel() - parts["r"].ravel()
        assert np.linalg.norm(res) < 1e-9 * np.linalg.norm(parts["r"])
    # same solution up to the constant null vector (0, 1, 1) that _shift_pressure fixes
    d = lam_mg - lam_bj
    d[0] -= d[0].mean()
    assert np.abs(d).max() < 1e-8 * np.abs(lam_bj).max()
    print(f"k={k} nx={nx}: trace CG iterations  multigrid {its_mg}  facet-block-Jacobi {its_bj}")
    assert its_mg <= 30 and its_mg < 0.5 * its_bj
