"""CPU execution of the tentative-velocity device kernels (csrc/hdg_flow.cuh, hdg_tent.cuh, hdg_advblock.cuh), compiled
with g++ through tests/host_kernels (test infrastructure; the engine has no CPU path), and a numpy port of the
host orchestration of `run_tentative_aug` (csrc/hdg_engine.cu): facet-multiplier formulation, Chebyshev /
facet-block-Jacobi sweeps on the facet Schur complement, BiCGStab on the augmented system.  Checks

* `k_fimpl` against the oracle's `f_impl` (`hdg_imex.py:313-331`),
* the solution of the tentative-velocity system `[M - a f_impl(.;Q*)] x = M b` (`hdg_imex.py:233-255`,
  `hdg_implicit.py:103-129`) against the oracle's sparse-direct solve,
* that the experimental cell-block advection preconditioner (knob ``tent_cellblock``) leaves the solution
  unchanged and cuts the iteration count with the real (inexact) sweeps, not only in the idealised model of
  tests/experiments/tent_precond_model.py.

The same kernels run on the GPU in tests/test_engine_flow_gpu.py / test_timesteppers_gpu.py."""
import ctypes
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from incompressibleeulerhdg_b200.mesh import UnitSquareMesh
from oracle.hdg_oracle import HDGOracle
from oracle.timesteppers import TaylorGreenOracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_kernels"))
import build as host_build  # noqa: E402

DP, IP, FP = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_float)


def dp(a):
    return None if a is None else a.ctypes.data_as(DP)


def ip(a):
    return a.ctypes.data_as(IP)


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    return host_build.build("tent_host.cpp", str(tmp_path_factory.mktemp("host_kernels")))


@pytest.fixture(scope="module")
def klib(tmp_path_factory):
    return host_build.build("krylov_host.cpp", str(tmp_path_factory.mktemp("host_kernels_krylov")))


def soa(Q):  # [nc, 2, nQ1] -> [2 nQ1][nc]
    return np.ascontiguousarray(Q.transpose(1, 2, 0).reshape(-1, Q.shape[0]))


def aos(Qs, nq1):  # [2 nQ1][nc] -> [nc, 2, nQ1]
    return np.ascontiguousarray(Qs.reshape(2, nq1, -1).transpose(2, 0, 1))


class HostTentative:
    """numpy port of tent_setup / tent_schur_solve / run_tentative_aug / bicgstab_loop (csrc/hdg_engine.cu)"""

    def __init__(self, lib, mesh, k, alpha=1.0, sweeps=8):
        self.lib, self.k, self.alpha, self.sweeps = lib, k, alpha, sweeps
        self.nc, self.nf = nc, nf = mesh.nc, mesh.nf
        self.nq1, self.nm = (k + 2) * (k + 3) // 2, k + 2
        i32 = lambda a: np.ascontiguousarray(np.asarray(a).T, dtype=np.int32)  # noqa: E731  AoS [n, m] -> SoA [m][n]
        self.xy = np.ascontiguousarray(np.asarray(mesh.cell_xy, dtype=np.float64).transpose(1, 2, 0).reshape(6, nc))
        self.cell_facet, self.cell_flip = i32(mesh.cell_facet), i32(mesh.cell_flip)
        self.facet_cell, self.facet_local = i32(mesh.facet_cell), i32(mesh.facet_local)
        self.nbr, self.nbr_e = np.zeros((3, nc), np.int32), np.zeros((3, nc), np.int32)
        self.tc, self.tcol, self.tbits = np.zeros((6, nf)), np.zeros((4, nf), np.int32), np.zeros(nf, np.int32)
        assert lib.th_setup(nc, nf, dp(self.xy), ip(self.cell_facet), ip(self.cell_flip), ip(self.facet_cell),
                            ip(self.facet_local), ip(self.nbr), ip(self.nbr_e), dp(self.tc), ip(self.tcol),
                            ip(self.tbits)) == 0
        self.d = np.zeros((self.nm, nf))
        self.lmax_global = self._power_iteration()
        lam = np.zeros(nc)
        assert lib.th_elem_bound(k, nc, dp(self.xy), dp(lam)) == 0
        self.lmax_elem = float(lam.max())
        self.lmax = max(self.lmax_global, 1.03 * self.lmax_elem / 1.1)  # tent_setup (csrc/hdg_engine.cu)
        self.cellblock = None
        self.sK = None
        self.tc0 = self.tc

    # -- kernels ---------------------------------------------------------------------------------------------
    def fimpl(self, upwind, Qstar, X, c0, c1, Z=None, alpha=None):
        Y = np.zeros_like(X)
        assert self.lib.th_fimpl(self.k, int(upwind), self.nc, dp(self.xy), ip(self.nbr), ip(self.nbr_e),
                                 ctypes.c_double(self.alpha if alpha is None else alpha), dp(Qstar), dp(X), dp(Z),
                                 ctypes.c_double(c0), ctypes.c_double(c1), dp(Y)) == 0
        return Y

    def sweep(self, inv_aalpha, rhs, x, cd, cr, zero, mode):
        xout = np.zeros((self.nm, self.nf))
        assert self.lib.th_sweep(self.k, self.nf, ip(self.facet_local), dp(self.tc), ip(self.tcol), ip(self.tbits),
                                 ctypes.c_double(inv_aalpha), dp(rhs), None, dp(x), dp(self.d), dp(xout),
                                 ctypes.c_double(cd), ctypes.c_double(cr), int(zero), int(mode)) == 0
        return xout

    def _power_iteration(self):
        x = np.random.default_rng(1).uniform(-0.5, 0.5, (self.nm, self.nf))
        lam = 2.0
        for _ in range(80):
            x = self.sweep(0.0, None, x, 0.0, 0.0, 0, 2)
            lam = np.linalg.norm(x)
            x /= lam
        return lam

    def schur_solve(self, inv_aalpha, t):
        a, b = self.lmax / 8.0, 1.1 * self.lmax  # cheb_coefs (csrc/hdg_mg.cuh)
        theta, delta = 0.5 * (b + a), 0.5 * (b - a)
        sigma = theta / delta
        rho = 1.0 / sigma
        coefs = [(0.0, 1.0 / theta)]
        for _ in range(1, self.sweeps):
            rho_new = 1.0 / (2.0 * sigma - rho)
            coefs.append((rho_new * rho, 2.0 * rho_new / delta))
            rho = rho_new
        x = np.zeros((self.nm, self.nf))
        for j, (cd, cr) in enumerate(coefs):
            x = self.sweep(inv_aalpha, t, x, cd, cr, j == 0, 0)
        return x

    def precond_x(self, inv_aalpha, in_x, in_mu):
        cm = np.zeros((3 * self.nm, self.nc))
        assert self.lib.th_moments(self.k, self.nc, dp(self.xy), ip(self.cell_flip), dp(in_x), dp(cm)) == 0
        t, nyx = np.zeros((self.nm, self.nf)), np.zeros((self.nm, self.nf))
        assert self.lib.th_trhs(self.k, self.nc, self.nf, dp(cm), ip(self.facet_cell), ip(self.facet_local), dp(in_mu),
                                dp(t), dp(nyx)) == 0
        return self.schur_solve(inv_aalpha, t), nyx

    def xhat(self, Y, mu, with_z=False):
        Xh = np.zeros_like(Y)
        Z = np.zeros_like(Y) if with_z else None
        assert self.lib.th_xhat_scaled(self.k, self.nc, self.nf, dp(self.xy), ip(self.cell_flip), ip(self.cell_facet),
                                       dp(Y), dp(mu), dp(Xh), 0, dp(self.sK), dp(Z)) == 0
        return (Xh, Z) if with_z else Xh

    def scaled_x(self, v):
        if self.cellblock is None:
            return v
        out = np.zeros_like(v)
        assert self.lib.th_advblock_apply(self.k, self.nc, self.cellblock.ctypes.data_as(FP), dp(v), dp(out)) == 0
        return out

    # -- run_tentative_aug -----------------------------------------------------------------------------------
    def make_op(self, Qstar, adt, upwind, use_cellblock, scaledx=True):
        """(op, split): op(v) -> (A_aug Phat^-1 v, [Phat^-1 v]_x) as in run_tentative_aug"""
        nq = 2 * self.nq1 * self.nc
        inv_aalpha = 1.0 / (adt * self.alpha)
        self.cellblock, self.sK, self.tc = None, None, self.tc0
        if use_cellblock:
            work = np.zeros((self.nq1 * self.nq1, self.nc))
            self.cellblock = np.zeros((self.nq1 * self.nq1, self.nc), np.float32)  # what the apply kernel reads
            sK = np.zeros(self.nc)
            assert self.lib.th_advblock_sk(self.k, int(upwind), self.nc, dp(self.xy), ip(self.nbr), dp(Qstar),
                                           ctypes.c_double(adt), dp(work), self.cellblock.ctypes.data_as(FP), dp(sK)) == 0
            if scaledx:  # scaled facet Schur complement (k_tent_scale_tc, csrc/hdg_tent.cuh)
                self.sK = sK
                self.tc = np.zeros_like(self.tc0)
                assert self.lib.th_scale_tc(self.nf, ip(self.facet_cell), dp(self.tc0), dp(sK), dp(self.tc)) == 0

        def split(v):
            return (np.ascontiguousarray(v[:nq].reshape(2 * self.nq1, self.nc)),
                    np.ascontiguousarray(v[nq:].reshape(self.nm, self.nf)))

        def op_xh(v):
            vx, vmu = split(v)
            in_x = self.scaled_x(vx)
            mu, nyx = self.precond_x(inv_aalpha, in_x, vmu)
            xh, z = self.xhat(in_x, mu, with_z=True)                              # z = xh + M^-1 N^T mu
            out_x = self.fimpl(upwind, Qstar, xh, 1.0, -adt, Z=z, alpha=0.0)      # z - a F0(xhat)
            out_mu = self.sweep(inv_aalpha, nyx, mu, 0.0, 0.0, 0, 1)              # N in_x - X mu
            return np.concatenate([out_x.ravel(), out_mu.ravel()]), xh

        return op_xh, split

    def fgmres(self, klib, Qstar, adt, upwind, b, rtol, use_cellblock, m=30, maxit=400, x0=None, scaledx=True):
        """replay of run_fgmres (csrc/hdg_engine.cu) with the k_gm_* kernels of csrc/hdg_krylov.cuh; returns
        (x, iterations, restart cycles)"""
        LONG = ctypes.c_long
        op_xh, _ = self.make_op(Qstar, adt, upwind, use_cellblock, scaledx)
        nq, nmu = 2 * self.nq1 * self.nc, self.nm * self.nf
        n = nq + nmu
        x = np.zeros((2 * self.nq1, self.nc)) if x0 is None else x0.copy()
        V, Z = np.zeros((m + 1, n)), np.zeros((m, nq))
        part, red, coef = np.zeros(m + 3), np.zeros(m + 8), np.zeros(m + 3)
        S_WW, S_NRM = m + 1, m + 2
        bb = float(b.ravel() @ b.ravel())
        H, cs, sn, g, y = np.zeros((m + 1, m)), np.zeros(m), np.zeros(m), np.zeros(m + 1), np.zeros(m)
        its = cycles = 0
        prev_rr, prev_est_conv = -1.0, False
        while True:
            cycles += 1
            t = self.fimpl(upwind, Qstar, x, 1.0, -adt)  # A x with the penalty
            V[0] = 0.0
            assert klib.kh_resid_norm(LONG(nq), dp(b), dp(t), dp(V[0]), dp(part)) == 0
            rr = red[0] = part[0]
            if rr <= rtol * rtol * bb or its >= maxit:
                return x, its, cycles
            if prev_est_conv and rr >= 0.25 * prev_rr and rr <= 1e4 * rtol * rtol * bb:  # round-off floor
                return x, its, cycles
            prev_rr, prev_est_conv = rr, False
            assert klib.kh_gm_scale(LONG(nq), dp(V[0]), dp(red), 0) == 0
            g[:] = 0.0
            g[0] = np.sqrt(rr)
            jj = 0
            for j in range(m):
                w, Z[j] = (lambda r: (r[0], r[1].ravel()))(op_xh(V[j]))
                V[j + 1] = w
                w = V[j + 1]
                for pas in range(2):
                    for i0 in range(0, j + 1, 8):
                        cnt = min(8, j + 1 - i0)
                        assert klib.kh_gm_dots(LONG(n), dp(w), dp(V), LONG(n), i0, cnt, S_WW if i0 == 0 else -1, dp(part)) == 0
                    red[:m + 3] = part  # k_part_finish with a single partial per slot
                    for i0 in range(0, j + 1, 8):
                        cnt = min(8, j + 1 - i0)
                        assert klib.kh_gm_axpy(LONG(n), dp(w), dp(V), LONG(n), i0, cnt, dp(red),
                                               S_NRM if i0 + cnt == j + 1 else -1, dp(part)) == 0
                    red[S_NRM] = part[S_NRM]
                    H[:j + 1, j] = (0.0 if pas == 0 else H[:j + 1, j]) + red[:j + 1]
                    ww, nrm2 = red[S_WW], red[S_NRM]
                    if pas == 0 and nrm2 >= 0.5 * ww:
                        break
                its += 1
                H[j + 1, j] = np.sqrt(max(nrm2, 0.0))
                for i in range(j):
                    a0, a1 = H[i, j], H[i + 1, j]
                    H[i, j], H[i + 1, j] = cs[i] * a0 + sn[i] * a1, -sn[i] * a0 + cs[i] * a1
                d = np.hypot(H[j, j], H[j + 1, j])
                cs[j], sn[j] = H[j, j] / d, H[j + 1, j] / d
                H[j, j], H[j + 1, j] = d, 0.0
                g[j + 1], g[j] = -sn[j] * g[j], cs[j] * g[j]
                jj = j + 1
                if g[j + 1] ** 2 <= rtol * rtol * bb:
                    prev_est_conv = True
                    break
                if its >= maxit:
                    break
                if j + 1 < m:
                    assert klib.kh_gm_scale(LONG(n), dp(w), dp(red), S_NRM) == 0
            y[:jj] = np.linalg.solve(np.triu(H[:jj, :jj]), g[:jj])
            coef[:jj] = y[:jj]
            xf = x.reshape(-1)
            for j0 in range(0, jj, 8):
                assert klib.kh_gm_update(LONG(nq), dp(xf), dp(Z), LONG(nq), j0, min(8, jj - j0), dp(coef)) == 0

    def solve(self, Qstar, adt, upwind, b, rtol, use_cellblock, maxit=400, x0=None, scaledx=True):
        nq, nmu = 2 * self.nq1 * self.nc, self.nm * self.nf
        inv_aalpha = 1.0 / (adt * self.alpha)
        op_xh, split = self.make_op(Qstar, adt, upwind, use_cellblock, scaledx)

        def op(v):
            return op_xh(v)[0]

        # BiCGStab on the augmented system; r0 = (b - A x0, 0) with mu0 = a alpha N x0 (zero guess: r0 = (b, 0))
        r0x = b if x0 is None else b - self.fimpl(upwind, Qstar, x0, 1.0, -adt)
        r = np.concatenate([r0x.ravel(), np.zeros(nmu)])
        bb = float(b.ravel() @ b.ravel())
        rhat, p, y = r.copy(), r.copy(), np.zeros(nq + nmu)
        rho, its = float(rhat @ r), 0
        while float(r @ r) > rtol * rtol * bb and its < maxit:
            its += 1
            v = op(p)
            al = rho / float(rhat @ v)
            s = r - al * v
            t = op(s)
            om = float(t @ s) / float(t @ t)
            y += al * p + om * s
            r = s - om * t
            rho_new = float(rhat @ r)
            if not np.isfinite(rho_new) or rho == 0.0 or om == 0.0:  # breakdown: report it as not converged
                its = maxit
                break
            p = r + (rho_new / rho) * (al / om) * (p - om * v)
            rho = rho_new
        yx, ymu = split(y)
        in_x = self.scaled_x(yx)
        mu, _ = self.precond_x(inv_aalpha, in_x, ymu)
        dx = self.xhat(in_x, mu)
        return (dx if x0 is None else x0 + dx), its


def _problem(k, nx, flux, cfl=0.32):
    mesh = UnitSquareMesh(nx, perturb=0.1)
    o = HDGOracle(mesh, k, alpha_penalty=1.0, flux=flux)
    prob = TaylorGreenOracle("exponential", 0.5)
    Q0 = o.interpolate_cell(lambda x, y: prob.Q_stationary(x, y), "Q")
    return mesh, o, Q0, o.project_bdm(Q0), cfl / nx


@pytest.mark.parametrize("flux", ["upwind", "centered"])
@pytest.mark.parametrize("k,nx", [(1, 4), (2, 3), (3, 2)])
def test_fimpl_kernel_on_the_host_matches_the_oracle(lib, k, nx, flux):
    mesh, o, Q0, Qs, _ = _problem(k, nx, flux)
    ht = HostTentative(lib, mesh, k)
    X = Q0 + 0.1 * np.random.default_rng(3).standard_normal(Q0.shape)
    got = aos(ht.fimpl(flux == "upwind", soa(Qs), soa(X), 0.0, 1.0), o.nQ1)       # M^-1 f_impl(., X; Q*)
    ref = o.f_impl_apply(X, Qs) / o.detJ[:, None, None]
    assert np.abs(got - ref).max() < 1e-11 * np.abs(ref).max()


@pytest.mark.parametrize("flux", ["upwind", "centered"])
@pytest.mark.parametrize("k,nx", [(1, 8), (2, 6)])
def test_tentative_solver_on_the_host(lib, k, nx, flux):
    mesh, o, Q0, Qs, adt = _problem(k, nx, flux)
    ht = HostTentative(lib, mesh, k)
    # spectrum of the block-Jacobi preconditioned facet Schur complement (hdg_tent.cuh): the element-wise bound holds
    # for every cell weighting, so it lies above the global power-iteration value
    assert 1.0 < ht.lmax_global <= ht.lmax_elem < 2.2
    rng = np.random.default_rng(11)
    b = Q0 + 0.01 * rng.standard_normal(Q0.shape)                                 # Riesz form: (I - a M^-1 f_impl) x = b
    M = sp.diags(np.repeat(o.detJ, o.nQ))
    x_ref = spla.spsolve((M - adt * o.f_impl_matrix(Qs)).tocsc(), M @ b.ravel()).reshape(b.shape)
    res = {}
    for cb in (False, True):
        x, its = ht.solve(soa(Qs), adt, flux == "upwind", soa(b), 1e-12, cb)
        err = np.abs(aos(x, o.nQ1) - x_ref).max() / np.abs(x_ref).max()
        res[cb] = (its, err)
        assert err < 1e-9, (cb, its, err)
    print(f"k={k} nx={nx} {flux}: BiCGStab iterations {res[False][0]} -> {res[True][0]} with the cell blocks")
    assert res[True][0] <= 0.75 * res[False][0]


@pytest.mark.parametrize("k,nx,cfl", [(2, 6, 0.32), (2, 6, 2.0), (1, 8, 4.0)])
def test_fgmres_fallback_on_the_host(lib, klib, k, nx, cfl):
    """the robust path of the tentative solve (run_fgmres; kernels k_gm_dots / k_gm_axpy / k_gm_scale / k_gm_update /
    k_resid_norm) reaches the oracle's sparse-direct solution at CFL numbers where BiCGStab needs many times the
    iterations or stalls (the reference's default dt = 0.04, src/driver.py:80-86, is that regime)"""
    mesh, o, Q0, Qs, adt = _problem(k, nx, "upwind", cfl)
    ht = HostTentative(lib, mesh, k, sweeps=4)
    b = Q0 + 0.01 * np.random.default_rng(11).standard_normal(Q0.shape)
    M = sp.diags(np.repeat(o.detJ, o.nQ))
    x_ref = spla.spsolve((M - adt * o.f_impl_matrix(Qs)).tocsc(), M @ b.ravel()).reshape(b.shape)
    x, its, cycles = ht.fgmres(klib, soa(Qs), adt, True, soa(b), 1e-12, True, m=40, maxit=1200)
    err = np.abs(aos(x, o.nQ1) - x_ref).max() / np.abs(x_ref).max()
    xb, its_b = ht.solve(soa(Qs), adt, True, soa(b), 1e-12, True, maxit=600)
    err_b = np.abs(aos(xb, o.nQ1) - x_ref).max() / np.abs(x_ref).max()
    xu, its_u, _ = ht.fgmres(klib, soa(Qs), adt, True, soa(b), 1e-12, True, m=40, maxit=1200, scaledx=False)
    print(f"k={k} nx={nx} cfl={cfl}: FGMRES(40) {its} iterations in {cycles} cycles, error {err:.1e} "
          f"({its_u} with the unscaled Schur complement); BiCGStab {its_b} iterations (2 applications each), "
          f"error {err_b:.1e}")
    assert err < 1e-9, (its, cycles, err)
    assert its <= its_u


@pytest.mark.parametrize("k", [1, 2, 3])
def test_float_instantiations_of_the_mixed_precision_kernels(lib, k):
    """the kernels the opt-in mixed-precision solver (run_tentative_mixed) instantiates with float -- k_fimpl in FP32
    arithmetic, moments / facet right-hand side / back-substitution / residual sweep with float storage -- against
    their FP64 versions on the same data: FP32 accuracy, nothing more and nothing less"""
    FP = ctypes.POINTER(ctypes.c_float)
    fp = lambda a: None if a is None else a.ctypes.data_as(FP)  # noqa: E731
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)  # noqa: E731
    mesh = UnitSquareMesh(4, perturb=0.15)
    ht = HostTentative(lib, mesh, k)
    rng = np.random.default_rng(k)
    X, Qs, Z = (rng.standard_normal((2 * ht.nq1, ht.nc)) for _ in range(3))
    mu = rng.standard_normal((ht.nm, ht.nf))
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())  # noqa: E731
    # operator
    Y64 = ht.fimpl(True, Qs, X, 1.0, -0.01, Z=Z, alpha=0.0)
    Y32 = np.zeros_like(X, dtype=np.float32)
    assert lib.th_fimpl32(k, 1, ht.nc, dp(ht.xy), ip(ht.nbr), ip(ht.nbr_e), ctypes.c_double(0.0), fp(f32(Qs)), fp(f32(X)),
                          fp(f32(Z)), ctypes.c_float(1.0), ctypes.c_float(-0.01), fp(Y32)) == 0
    assert rel(Y32, Y64) < 2e-5
    # moments, facet right-hand side
    cm64 = np.zeros((3 * ht.nm, ht.nc))
    assert lib.th_moments(k, ht.nc, dp(ht.xy), ip(ht.cell_flip), dp(X), dp(cm64)) == 0
    cm32 = np.zeros_like(cm64, dtype=np.float32)
    assert lib.th_moments32(k, ht.nc, dp(ht.xy), ip(ht.cell_flip), fp(f32(X)), fp(cm32)) == 0
    assert rel(cm32, cm64) < 2e-6
    t64, n64 = np.zeros((ht.nm, ht.nf)), np.zeros((ht.nm, ht.nf))
    assert lib.th_trhs(k, ht.nc, ht.nf, dp(cm64), ip(ht.facet_cell), ip(ht.facet_local), dp(mu), dp(t64), dp(n64)) == 0
    t32, n32 = np.zeros_like(t64, dtype=np.float32), np.zeros_like(t64, dtype=np.float32)
    assert lib.th_trhs32(k, ht.nc, ht.nf, fp(f32(cm64)), ip(ht.facet_cell), ip(ht.facet_local), fp(f32(mu)), fp(t32),
                         fp(n32)) == 0
    assert rel(t32, t64) < 2e-6 and rel(n32, n64) < 2e-6
    # back-substitution of the multiplier with the scaled Schur complement, residual row of the operator
    sK = 0.5 + rng.random(ht.nc)
    ht.sK = sK
    Xh64, Z64 = ht.xhat(X, mu, with_z=True)
    Xh32, Z32 = np.zeros_like(X, dtype=np.float32), np.zeros_like(X, dtype=np.float32)
    assert lib.th_xhat32(k, ht.nc, ht.nf, dp(ht.xy), ip(ht.cell_flip), ip(ht.cell_facet), fp(f32(X)), fp(f32(mu)),
                         fp(Xh32), dp(sK), fp(Z32)) == 0
    assert rel(Xh32, Xh64) < 2e-6 and rel(Z32, Z64) < 2e-6
    r64 = ht.sweep(3.0, n64, mu, 0.0, 0.0, 0, 1)
    r32 = np.zeros_like(r64, dtype=np.float32)
    assert lib.th_sweep_f(k, ht.nf, ip(ht.facet_local), dp(ht.tc), ip(ht.tcol), ip(ht.tbits), ctypes.c_double(3.0),
                          fp(f32(n64)), fp(f32(mu)), fp(r32), 1) == 0
    assert rel(r32, r64) < 5e-6


@pytest.mark.parametrize("k", [1, 2, 3, 4])
@pytest.mark.parametrize("upwind", [True, False])
def test_operator_with_tabulated_advecting_velocity_is_the_same_operator(lib, k, upwind):
    """k_fimpl_pre + k_fimpl_q (the operator of the Krylov iterations: everything derived from the fixed Q* is
    tabulated once per solve) against k_fimpl: the same floating-point operations in the same order, hence identical"""
    mesh = UnitSquareMesh(4, perturb=0.15)
    ht = HostTentative(lib, mesh, k)
    rng = np.random.default_rng(10 + k)
    X, Qs, Z = (rng.standard_normal((2 * ht.nq1, ht.nc)) for _ in range(3))
    npre = ctypes.c_int(0)
    assert lib.th_fimpl_pre(k, ht.nc, dp(ht.xy), dp(Qs), None, ctypes.byref(npre)) == 0
    pre = np.zeros((npre.value, ht.nc))
    assert lib.th_fimpl_pre(k, ht.nc, dp(ht.xy), dp(Qs), dp(pre), ctypes.byref(npre)) == 0
    for alpha, Zarg in ((0.0, Z), (1.0, None)):
        Y = ht.fimpl(upwind, Qs, X, 1.0, -0.37, Z=Zarg, alpha=alpha)
        Yq = np.zeros_like(X)
        assert lib.th_fimpl_q(k, int(upwind), ht.nc, dp(ht.xy), ip(ht.nbr), ip(ht.nbr_e), ctypes.c_double(alpha), dp(pre),
                              dp(X), dp(Zarg), ctypes.c_double(1.0), ctypes.c_double(-0.37), dp(Yq)) == 0
        assert np.abs(Yq - Y).max() <= 1e-13 * np.abs(Y).max()


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_gram_blocks_do_not_depend_on_the_local_facet(lib, k):
    """what k_tent_sweep / k_tent_sweep32 rely on to run without a switch over the local facet index (csrc/hdg_tent.cuh):
    GG(e, (e + j) % 3) = GG(0, j) for every local facet e, and the facet's own block GG(e, e) is diagonal"""
    nm = k + 2
    G = np.zeros((3, 3, nm, nm))
    out = ctypes.c_double()
    for e in range(3):
        for f in range(3):
            for j in range(nm):
                for l in range(nm):  # noqa: E741
                    assert lib.th_gram(k, e, f, j, l, ctypes.byref(out)) == 0
                    G[e, f, j, l] = out.value
    for e in range(3):
        for jj in range(3):
            assert np.abs(G[e, (e + jj) % 3] - G[0, jj]).max() <= 4e-16 * np.abs(G[0, jj]).max()
        assert np.abs(G[e, e] - np.diag(np.diag(G[e, e]))).max() == 0.0
        assert np.all(np.diag(G[e, e]) > 0.0)


@pytest.mark.parametrize("k", [1, 2, 3, 4])
@pytest.mark.parametrize("upwind", [True, False])
def test_component_split_operator_is_the_penalty_free_operator(lib, k, upwind):
    """k_fimpl_c (one thread per (cell, component), Q* from the table of k_fimpl_pre; the operator of the augmented
    Krylov iteration, where the penalty lives in the multiplier rows) against k_fimpl with alpha = 0, with and without
    a separate Z, on a mesh whose cell count is not a multiple of 16"""
    for nx in (3, 5):
        mesh = UnitSquareMesh(nx, perturb=0.15)
        ht = HostTentative(lib, mesh, k)
        rng = np.random.default_rng(20 + k)
        X, Qs, Z = (rng.standard_normal((2 * ht.nq1, ht.nc)) for _ in range(3))
        npre = ctypes.c_int(0)
        assert lib.th_fimpl_pre(k, ht.nc, dp(ht.xy), dp(Qs), None, ctypes.byref(npre)) == 0
        pre = np.zeros((npre.value, ht.nc))
        assert lib.th_fimpl_pre(k, ht.nc, dp(ht.xy), dp(Qs), dp(pre), ctypes.byref(npre)) == 0
        for Zarg in (Z, None):
            Y = ht.fimpl(upwind, Qs, X, 1.0, -0.37, Z=Zarg, alpha=0.0)
            Yc = np.full_like(X, np.nan)
            assert lib.th_fimpl_c(k, int(upwind), ht.nc, dp(ht.xy), ip(ht.nbr), ip(ht.nbr_e), dp(pre), dp(X), dp(Zarg),
                                  ctypes.c_double(1.0), ctypes.c_double(-0.37), dp(Yc)) == 0
            assert np.abs(Yc - Y).max() <= 1e-13 * np.abs(Y).max()
