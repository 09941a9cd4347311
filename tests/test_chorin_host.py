"""Capstone of the CPU execution of the device code: two complete Chorin timesteps
(`hdg_implicit.py:92-190`: BDM projection, tentative velocity, weak divergence, static condensation + forward
elimination, multigrid-preconditioned trace solve, back-substitution, velocity / pressure update) computed by the
engine's *device kernels*, compiled with g++ and strung together by numpy ports of the host orchestration, against
the oracle's timestepper (sparse-direct solves).  What stays GPU-only are the vector kernels of CG / BiCGStab and the
orchestration in hdg_engine.cu itself (tests/test_timesteppers_gpu.py).  Test infrastructure: the engine has no
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
from test_mg_host import HostGTMG  # noqa: E402
from test_poisson_host import HostMesh, dp, ip, rel  # noqa: E402
from test_tent_host import HostTentative, aos, soa  # noqa: E402

cd = ctypes.c_double


@pytest.fixture(scope="module")
def libs(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host_kernels"))
    return {n: host_build.build(n + "_host.cpp", out) for n in ("poisson", "mg", "flow", "tent")}


class HostChorin:
    """IncompressibleEulerHDGImplicit.step on the host-compiled kernels (Riesz-form fields in the engine's SoA layout)"""

    def __init__(self, libs, mesh, k, dt, flux):
        self.libs, self.mesh, self.k, self.dt, self.upwind = libs, mesh, k, dt, flux == "upwind"
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
        lam, its = self.mg.pcg(np.ascontiguousarray(b), self.mg.apply, rtol=1e-13)
        u, phi = np.zeros((2 * self.nq1, nc)), np.zeros((self.np_, nc))
        assert lp.ph_back(k, nc, nf, dp(hm.xy), ip(hm.cell_flip), ip(hm.cell_facet), cd(1.0), None, dp(Rp),
                          dp(np.ascontiguousarray(lam)), dp(u), dp(phi)) == 0
        return u, phi, lam, its

    def step(self, Q, f):
        dt, hm, nc = self.dt, self.hm, self.mesh.nc
        Qstar = self.project_bdm(Q)                                                    # :98
        rhs = Q + dt * f                                                               # :126 in Riesz form
        Qt, its_t = self.tent.solve(Qstar, dt, self.upwind, rhs, 1e-13, False)         # :129
        Rp = np.zeros((self.np_, nc))
        assert self.libs["flow"].fh_weak_div(self.k, nc, dp(hm.xy), ip(self.tent.nbr), ip(self.tent.nbr_e), dp(Qt),
                                             cd(-1.0 / dt), 0, dp(Rp)) == 0            # :145
        u, phi, lam, its_p = self.poisson_apply(Rp)                                    # :146
        return Qt + dt * u, phi, (its_t, its_p)                                        # :150, :189


@pytest.mark.parametrize("k,nx,flux", [(1, 6, "upwind"), (2, 4, "upwind"), (2, 4, "centered"), (3, 3, "centered")])
def test_two_chorin_steps_from_device_kernels(libs, k, nx, flux):
    mesh, dt = UnitSquareMesh(nx, perturb=0.1), 0.02
    orc = ChorinOracle(mesh, k, dt, flux=flux)
    prob = TaylorGreenOracle("exponential", 0.5)
    Qo, po = orc.initial_state(prob)
    hc = HostChorin(libs, mesh, k, dt, flux)
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
