"""Capstone of the CPU execution of the device code: two complete Chorin timesteps
(`hdg_implicit.py:92-190`: BDM projection, tentative velocity, weak divergence, static condensation + forward
elimination, multigrid-preconditioned trace solve, back-substitution, velocity / pressure update) computed by the
engine's *device kernels*, compiled with g++ and strung together by numpy ports of the host orchestration, against
the oracle's timestepper (sparse-direct solves).  With `krylov_kernels` the CG / BiCGStab vector kernels are the engine's
too (tests/test_krylov_host.py), so that every arithmetic operation of the step is device code; what stays GPU-only is
the orchestration in hdg_engine.cu itself (tests/test_timesteppers_gpu.py).  Test infrastructure: the engine has no
CPU path."""
import ctypes
import os
import sys

import numpy as np
import pytest

from incompressibleeulerhdg_b200.mesh import UnitSquareMesh
from oracle.timesteppers import ChorinOracle, TaylorGreenOracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_kernels"))
import build as host_build  # noqa: E402
from test_krylov_host import bicgstab_kernels, pcg_mg_kernels  # noqa: E402
from test_mg_host import HostGTMG  # noqa: E402
from test_poisson_host import HostMesh, dp, ip, rel  # noqa: E402
from test_tent_host import HostTentative, aos, soa  # noqa: E402

cd = ctypes.c_double


@pytest.fixture(scope="module")
def libs(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host_kernels"))
    return {n: host_build.build(n + "_host.cpp", out) for n in ("poisson", "mg", "flow", "tent", "krylov")}


class HostChorin:
    """IncompressibleEulerHDGImplicit.step on the host-compiled kernels (Riesz-form fields in the engine's SoA layout)"""

    def __init__(self, libs, mesh, k, dt, flux, krylov_kernels=False):
        """krylov_kernels: run CG / BiCGStab with the engine's vector kernels (launch sequences of run_pcg_mg and
        bicgstab_loop) instead of the numpy loops of test_mg_host.py / test_tent_host.py"""
        self.libs, self.mesh, self.k, self.dt, self.upwind = libs, mesh, k, dt, flux == "upwind"
        self.krylov_kernels = krylov_kernels
        self.hm = HostMesh(mesh)
        self.tent = HostTentative(libs["tent"], mesh, k)
        self.mg = HostGTMG((libs["poisson"], libs["mg"]), mesh, k)
        self.nq1, self.np_, self.nl1 = (k + 2) * (k + 3) // 2, (k + 1) * (k + 2) // 2, k + 1

    def project_bdm(self, Q):
        hm, nf = self.hm, self.mesh.nf
        fm, out = np.zeros((2 * (self.k + 2), nf)), np.zeros_like(Q)
        assert self.libs["flow"].fh_project_bdm(self.k, hm.nc, nf, dp(hm.xy), ip(hm.cell_facet), ip(hm.facet_cell), dp(Q),
                                                dp(fm), dp(out)) == 0
        return out

    def poisson_apply(self, Rp):
        """hdg_poisson_apply_dev with a pressure right-hand side only: forward elimination, trace right-hand side,
        MG-CG on P = -S, back-substitution (the pressure shift is left to the caller)"""
        hm, k, nc, nf, nl1 = self.hm, self.k, self.mesh.nc, self.mesh.nf, self.nl1
        lp = self.libs["poisson"]
        gK = np.zeros((3 * nl1, nc))
        assert lp.ph_forward(k, nc, dp(hm.xy), ip(hm.cell_flip), cd(1.0), None, dp(Rp), dp(gK)) == 0
        # k_trace_rhs (csrc/hdg_engine.cu): b = sum over the adjacent cells of gK  (= -(R_l - sum gK) with R_l = 0)
        fc, fl = self.mesh.facet_cell, self.mesh.facet_local
        g3 = gK.reshape(3, nl1, nc)
        b = g3[fl[:, 0], :, fc[:, 0]].T.copy()
        interior = fc[:, 1] >= 0
        b[:, interior] += g3[fl[interior, 1], :, fc[interior, 1]].T
        if self.krylov_kernels:
            lk = self.libs["krylov"]
            b2, pm = np.zeros((nl1, nf)), np.zeros(1)
            assert lk.kh_trace_rhs(k, nc, nf, dp(gK), None, ip(hm.facet_cell), ip(hm.facet_local), dp(b2), dp(pm)) == 0
            assert np.array_equal(b2, b)
            lam, (_, _, its, done) = pcg_mg_kernels(lk, self.mg, b2, pm[0], 1e-13)
            assert done == 1
        else:
            lam, its = self.mg.pcg(np.ascontiguousarray(b), self.mg.apply, rtol=1e-13)
        u, phi = np.zeros((2 * self.nq1, nc)), np.zeros((self.np_, nc))
        assert lp.ph_back(k, nc, nf, dp(hm.xy), ip(hm.cell_flip), ip(hm.cell_facet), cd(1.0), None, dp(Rp),
                          dp(np.ascontiguousarray(lam)), dp(u), dp(phi)) == 0
        return u, phi, lam, its

    def tentative_with_kernels(self, Qstar, rhs):
        """run_tentative_aug with bicgstab_loop on the engine's BiCGStab kernels (zero initial guess)"""
        ht, adt = self.tent, self.dt
        nq, nmu = 2 * ht.nq1 * ht.nc, ht.nm * ht.nf
        inv_aalpha = 1.0 / (adt * ht.alpha)

        def split(vec):
            return (np.ascontiguousarray(vec[:nq].reshape(2 * ht.nq1, ht.nc)),
                    np.ascontiguousarray(vec[nq:].reshape(ht.nm, ht.nf)))

        def op(vec):
            vx, vmu = split(vec)
            mu, nyx = ht.precond_x(inv_aalpha, vx, vmu)
            xh = ht.xhat(vx, mu)
            out_x = ht.fimpl(self.upwind, Qstar, xh, 1.0, -adt, Z=vx, alpha=0.0)
            out_mu = ht.sweep(inv_aalpha, nyx, mu, 0.0, 0.0, 0, 1)
            return np.concatenate([out_x.ravel(), out_mu.ravel()])

        r0 = np.concatenate([rhs.ravel(), np.zeros(nmu)])
        y, its, done = bicgstab_kernels(self.libs["krylov"], op, nq + nmu, r0, float(rhs.ravel() @ rhs.ravel()), 1e-13, 400)
        assert done == 1
        yx, ymu = split(y)
        mu, _ = ht.precond_x(inv_aalpha, yx, ymu)
        return ht.xhat(yx, mu), its

    def step(self, Q, f):
        dt, hm, nc = self.dt, self.hm, self.mesh.nc
        Qstar = self.project_bdm(Q)                                                    # :98
        rhs = Q + dt * f                                                               # :126 in Riesz form
        if self.krylov_kernels:
            Qt, its_t = self.tentative_with_kernels(Qstar, rhs)                        # :129
        else:
            Qt, its_t = self.tent.solve(Qstar, dt, self.upwind, rhs, 1e-13, False)     # :129
        Rp = np.zeros((self.np_, nc))
        assert self.libs["flow"].fh_weak_div(self.k, nc, dp(hm.xy), ip(self.tent.nbr), ip(self.tent.nbr_e), dp(Qt),
                                             cd(-1.0 / dt), 0, dp(Rp)) == 0            # :145
        u, phi, lam, its_p = self.poisson_apply(Rp)                                    # :146
        return Qt + dt * u, phi, (its_t, its_p)                                        # :150, :189


@pytest.mark.parametrize("k,nx,flux,krylov_kernels", [(1, 6, "upwind", False), (2, 4, "upwind", True),
                                                      (2, 4, "centered", False), (3, 3, "centered", True)])
def test_two_chorin_steps_from_device_kernels(libs, k, nx, flux, krylov_kernels):
    mesh, dt = UnitSquareMesh(nx, perturb=0.1), 0.02
    orc = ChorinOracle(mesh, k, dt, flux=flux)
    prob = TaylorGreenOracle("exponential", 0.5)
    Qo, po = orc.initial_state(prob)
    hc = HostChorin(libs, mesh, k, dt, flux, krylov_kernels=krylov_kernels)
    Q = soa(Qo)
    for step in range(2):
        f_fun = prob.f_rhs(step * dt)
        Qo, po = orc.step(Qo, po, f_fun)
        Q, phi, its = hc.step(Q, soa(orc.interp_Q(f_fun)))
        p = np.ascontiguousarray(phi.T)
        p = p - orc.o.integral_p(p) / mesh.volume * orc.o.const_p()                    # _shift_pressure, :189-190
        eQ, ep = rel(aos(Q, orc.o.nQ1), Qo), rel(p, po)
        print(f"k={k} {flux} step {step}: velocity {eQ:.1e} pressure {ep:.1e}  BiCGStab / MG-CG iterations {its}")
        assert eQ < 1e-9 and ep < 1e-9
