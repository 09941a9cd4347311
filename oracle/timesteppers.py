"""CPU oracle for the timestep loops  --  TEST INFRASTRUCTURE ONLY (see hdg_oracle.py header).

PARITY UNPINNED (no reference tests / golden vectors exist; Firedrake is not installable here).

Restates, with sparse *direct* solves everywhere (so that it is the "exact" answer the
iterative GPU engine must reach):

* `IncompressibleEulerHDGImplicit.solve`  src/timesteppers/hdg_implicit.py:52-197
  (Chorin projection :101-150 and the fully implicit monolithic step :151-186)
* `IncompressibleEulerHDGIMEX.solve`      src/timesteppers/hdg_imex.py:505-660 with
  `_residual` :367-391, `_final_residual` :393-413, the Richardson/projection stage :570-599,
  the unsplit stage :600-620, final stage :624 and pressure reconstruction :629-637
* the five tableaux                        hdg_imex.py:702-729,766-799,836-879,916-949,986-1038
  including the quirk that ARS3(4,4,3)._b_impl has six entries (:874) of which `_final_residual`
  uses indices 1..4.
* Taylor-Green problem                     src/model_problems.py:38-105

Fields are AoS modal coefficient arrays as in hdg_oracle.py.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .hdg_oracle import HDGOracle

__all__ = ["TaylorGreenOracle", "ChorinOracle", "IMEXOracle", "TABLEAUX"]


# ------------------------------------------------------------------------------------------------
# model problem (model_problems.py:38-105)
# ------------------------------------------------------------------------------------------------
class TaylorGreenOracle:
    def __init__(self, forcing="exponential", kappa=0.5):
        assert forcing in ("exponential", "constant")
        self.forcing, self.kappa = forcing, kappa

    @staticmethod
    def Q_stationary(x, y):
        S, Cc = np.sin, np.cos
        return (-Cc((x - 0.5) * np.pi) * S((y - 0.5) * np.pi), S((x - 0.5) * np.pi) * Cc((y - 0.5) * np.pi))

    @staticmethod
    def p_stationary(x, y):
        return (np.sin((x - 0.5) * np.pi) ** 2 + np.sin((y - 0.5) * np.pi) ** 2) / 2

    def psi(self, t):
        return np.exp(-self.kappa * t) if self.forcing == "exponential" else 1.0 - self.kappa * t

    def dpsi(self, t):
        return -self.kappa * np.exp(-self.kappa * t) if self.forcing == "exponential" else -self.kappa

    def f_rhs(self, t):
        """forcing f(t) = Psi'(t) Q_s as a callable of (x, y)   (model_problems.py:71-80)"""
        c = self.dpsi(t)
        return lambda x, y: tuple(c * v for v in self.Q_stationary(x, y))


# ------------------------------------------------------------------------------------------------
# shared pieces
# ------------------------------------------------------------------------------------------------
class _Base:
    def __init__(self, mesh, degree, dt, flux="upwind", tau=1.0, alpha=1.0, nq_facet=None):
        self.o = HDGOracle(mesh, degree, tau=tau, alpha_penalty=alpha, flux=flux, nq_facet=nq_facet)
        self.mesh, self.k, self.dt = mesh, degree, dt
        self._lu_cache = {}

    # -- interpolation (Function.interpolate) -----------------------------------------------------
    def interp_Q(self, fun):
        return self.o.interpolate_cell(fun, "Q")

    def interp_p(self, fun):
        return self.o.interpolate_cell(fun, "p")

    def initial_state(self, problem):
        o = self.o
        Q = self.interp_Q(problem.Q_stationary)
        p = self.interp_p(problem.p_stationary)
        p = p - o.integral_p(p) / self.mesh.volume * o.const_p()
        return Q, p

    # -- operators -----------------------------------------------------------------------------------
    def mass(self, Q):
        return self.o.mass_Q(Q)

    def tentative_matrix(self, Qstar, adt):
        """M - a dt f_impl(.,.;Q*)   (hdg_imex.py:233-235; hdg_implicit.py:103-125)"""
        o = self.o
        n = self.mesh.nc * o.nQ
        M = sp.diags(np.repeat(o.detJ, o.nQ))
        return (M - adt * o.f_impl_matrix(Qstar)).tocsc()

    def monolithic_implicit(self, Qstar, adt):
        """[[M - adt F, -adt B^T, adt E^T],[B, T, -tau F^T],[E, tau F, -tau G]]
        (hdg_imex.py:602-610; hdg_implicit.py:153-183)"""
        o = self.o
        K, (offp, offl, N) = o.assemble_monolithic()
        K = K.tolil()
        nq = offp
        Kc = K.tocsr()
        top = Kc[:nq]
        # scale the pressure-gradient columns of the w-rows by adt and add -adt F
        Mdiag = sp.diags(np.repeat(o.detJ, o.nQ))
        top_uu = Mdiag - adt * o.f_impl_matrix(Qstar)
        top_rest = adt * top[:, nq:]
        new_top = sp.hstack([top_uu, top_rest])
        return sp.vstack([new_top, Kc[nq:]]).tocsr(), (offp, offl, N)

    def solve_pinned(self, K, rhs, offl, N):
        keep = np.ones(N, dtype=bool)
        keep[offl] = False
        x = np.zeros(N)
        x[keep] = spla.splu(K[keep][:, keep].tocsc()).solve(rhs[keep])
        return x


# ------------------------------------------------------------------------------------------------
# hdg_implicit.py
# ------------------------------------------------------------------------------------------------
class ChorinOracle(_Base):
    """`IncompressibleEulerHDGImplicit` (hdg_implicit.py:10-197)"""

    def __init__(self, mesh, degree, dt, flux="upwind", use_projection_method=True, **kw):
        super().__init__(mesh, degree, dt, flux, **kw)
        self.use_projection_method = use_projection_method

    def step(self, Q, p, f_fun):
        o, dt = self.o, self.dt
        nc = self.mesh.nc
        Qstar = o.project_bdm(Q)  # :98
        f = self.interp_Q(f_fun)  # :100
        rhs = self.mass(Q) + dt * self.mass(f)  # :126 / :182
        if self.use_projection_method:
            A = self.tentative_matrix(Qstar, dt)
            Qt = spla.splu(A).solve(rhs.ravel()).reshape(nc, 2, o.nQ1)  # :129
            Rp = -(1.0 / dt) * o.cell_divergence(Qt)  # :145
            u, phi, lam = o.solve_condensed(np.zeros_like(Q), Rp, np.zeros((self.mesh.nf, o.nl1)))  # :146
            Qn = Qt + dt * u  # :150
            self.last = dict(Qstar=Qstar, Qt=Qt, u=u, phi=phi, lam=lam, Rp=Rp)
        else:
            K, (offp, offl, N) = self.monolithic_implicit(Qstar, dt)
            R = np.concatenate([rhs.ravel(), np.zeros(N - offp)])
            x = self.solve_pinned(K, R, offl, N)  # :185
            Qn = x[:offp].reshape(nc, 2, o.nQ1)
            phi = x[offp:offl].reshape(nc, o.np_)
            lam = x[offl:].reshape(self.mesh.nf, o.nl1)
            self.last = dict(Qstar=Qstar, lam=lam)
        pn = phi - o.integral_p(phi) / self.mesh.volume * o.const_p()  # :189-190
        return Qn, pn

    def solve(self, problem, T_final, warmup=False, q_initial=None):
        """with `q_initial` (a callable) the passive tracer is advected as in hdg_implicit.py:73-77,
        93-96,192-193 -- explicit Euler with the CG-projected velocity of the *old* time level -- and
        the result is kept in ``self.q_tracer``"""
        nt = 1 if warmup else int(np.round(T_final / self.dt))
        Q, p = self.initial_state(problem)
        q = None
        if q_initial is not None:
            from .tracer import TracerOracle

            tr = TracerOracle(self.o)
            q = self.interp_p(q_initial)
        for k in range(nt):
            if q is not None:
                u_cg = tr.project_cg(Q)  # :93-96 (project runs when the form is built)
            Q, p = self.step(Q, p, problem.f_rhs(k * self.dt))
            if q is not None:
                q = q + self.dt * tr.advection(q, u_cg)  # :192-193
        self.q_tracer = q
        return Q, p


# ------------------------------------------------------------------------------------------------
# hdg_imex.py
# ------------------------------------------------------------------------------------------------
def _ars2():
    g = 1 - 1 / np.sqrt(2)
    d = -2 / 3 * np.sqrt(2)
    return dict(nstages=3, a_expl=[[0, 0, 0], [g, 0, 0], [d, 1 - d, 0]], a_impl=[[0, 0, 0], [0, g, 0], [0, 1 - g, g]],
                b_expl=[0, 1 - g, g], b_impl=[0, 1 - g, g], c_expl=[0, g, 1])


def _ssp3():
    al, be, et = 0.24169426078821, 0.06042356519705, 0.12915286960590
    de = 1 / 2 - al - be - et
    return dict(nstages=4, a_expl=[[0, 0, 0, 0], [0, 0, 0, 0], [0, 1, 0, 0], [0, 1 / 4, 1 / 4, 0]],
                a_impl=[[al, 0, 0, 0], [-al, al, 0, 0], [0, 1 - al, al, 0], [be, et, de, al]],
                b_expl=[0, 1 / 6, 1 / 6, 2 / 3], b_impl=[0, 1 / 6, 1 / 6, 2 / 3], c_expl=[0, 0, 1, 1 / 2])


#: hdg_imex.py:702-729, 766-799, 836-879, 916-949, 986-1038
TABLEAUX = {
    "imex_implicit": dict(nstages=2, a_expl=[[0, 0], [1, 0]], a_impl=[[0, 0], [0, 1]], b_expl=[1, 0], b_impl=[0, 1],
                          c_expl=[0, 1]),
    "imex_ars2_232": _ars2(),
    "imex_ars3_443": dict(nstages=5,
                          a_expl=[[0, 0, 0, 0, 0], [1 / 2, 0, 0, 0, 0], [11 / 18, 1 / 18, 0, 0, 0],
                                  [5 / 6, -5 / 6, 1 / 2, 0, 0], [1 / 4, 7 / 4, 3 / 4, -7 / 4, 0]],
                          a_impl=[[0, 0, 0, 0, 0], [0, 1 / 2, 0, 0, 0], [0, 1 / 6, 1 / 2, 0, 0],
                                  [0, -1 / 2, 1 / 2, 1 / 2, 0], [0, 3 / 2, -3 / 2, 1 / 2, 1 / 2]],
                          b_expl=[1 / 4, 7 / 4, 3 / 4, -7 / 4, 0], b_impl=[0, 3 / 2, -3, 2, 1 / 2, 1 / 2],
                          c_expl=[0, 1 / 2, 2 / 3, 1 / 2, 1]),
    "imex_ssp2_332": dict(nstages=3, a_expl=[[0, 0, 0], [1 / 2, 0, 0], [1 / 2, 1 / 2, 0]],
                          a_impl=[[1 / 4, 0, 0], [0, 1 / 4, 0], [1 / 3, 1 / 3, 1 / 3]], b_expl=[1 / 3, 1 / 3, 1 / 3],
                          b_impl=[1 / 3, 1 / 3, 1 / 3], c_expl=[0, 1, 1 / 2]),
    "imex_ssp3_433": _ssp3(),
}


class IMEXOracle(_Base):
    """`IncompressibleEulerHDGIMEX` (hdg_imex.py:22-660)"""

    def __init__(self, mesh, degree, dt, tableau="imex_ssp2_332", flux="upwind", use_projection_method=True,
                 n_richardson=2, **kw):
        super().__init__(mesh, degree, dt, flux, **kw)
        t = TABLEAUX[tableau]
        self.nstages = t["nstages"]
        self.a_expl = np.asarray(t["a_expl"], dtype=float)
        self.a_impl = np.asarray(t["a_impl"], dtype=float)
        self.b_expl = np.asarray(t["b_expl"], dtype=float)
        self.b_impl = np.asarray(t["b_impl"], dtype=float)
        self.c_expl = np.asarray(t["c_expl"], dtype=float)
        self.use_projection_method = use_projection_method
        self.n_richardson = n_richardson
        o = self.o
        nc, nf = mesh.nc, mesh.nf
        zQ = lambda: np.zeros((nc, 2, o.nQ1))
        # persistent stage state (hdg_imex.py:72-88): never reset between timesteps
        self.stage = [dict(Q=zQ(), p=np.zeros((nc, o.np_)), l=np.zeros((nf, o.nl1))) for _ in range(self.nstages)]
        self.b_rhs = [zQ() for _ in range(self.nstages)]
        self.tracer, self.q_tracer = None, None

    # residual recursions as dual vectors -----------------------------------------------------------
    def residual(self, i):
        assert 0 < i < self.nstages
        r = self.mass(self.stage[0]["Q"])
        for j in range(1, i):
            if self.a_impl[i, j] != 0:
                r = r + self.a_impl[i, j] / self.a_impl[j, j] * (self.mass(self.stage[j]["Q"]) - self.residual(j))
        for j in range(i):
            if self.a_expl[i, j] != 0:
                r = r + self.dt * self.a_expl[i, j] * self.mass(self.b_rhs[j])
        return r

    def final_residual(self):
        r = self.mass(self.stage[0]["Q"])
        for i in range(1, self.nstages):
            if self.b_impl[i] != 0:
                r = r + self.b_impl[i] / self.a_impl[i, i] * (self.mass(self.stage[i]["Q"]) - self.residual(i))
        for i in range(self.nstages):
            if self.b_expl[i] != 0:
                r = r + self.dt * self.b_expl[i] * self.mass(self.b_rhs[i])
        return r

    def step(self, cur, problem, tn):
        o, dt = self.o, self.dt
        nc, nf = self.mesh.nc, self.mesh.nf
        zQ, zp, zl = np.zeros((nc, 2, o.nQ1)), np.zeros((nc, o.np_)), np.zeros((nf, o.nl1))
        for i in range(self.nstages):  # :554-557
            self.b_rhs[i] = self.interp_Q(problem.f_rhs(tn + self.c_expl[i] * dt))
        self.stage[0] = dict(Q=cur["Q"].copy(), p=cur["p"].copy(), l=cur["l"].copy())  # :558
        for i in range(1, self.nstages):
            st = self.stage[i]
            a = self.a_impl[i, i]
            Qstar = o.project_bdm(self.stage[i - 1]["Q"])  # :564-567
            if self.use_projection_method:
                A = spla.splu(self.tentative_matrix(Qstar, a * dt))
                for _ in range(self.n_richardson):  # :570
                    rhs = (self.residual(i) - self.mass(st["Q"])
                           + a * dt * (o.f_impl_apply(st["Q"], Qstar) + o.pressure_gradient(st["p"], st["l"])))  # :239-247
                    Qt = A.solve(rhs.ravel()).reshape(nc, 2, o.nQ1)  # :572
                    Rp = -1.0 / (a * dt) * o.weak_divergence(Qt)  # :177-179
                    u, phi, lam = o.solve_condensed(zQ, Rp, zl)  # :575 + _shift_pressure(update) :579
                    st["Q"] = st["Q"] + Qt + a * dt * u  # :580-587
                    st["p"] = st["p"] + phi
                    st["l"] = st["l"] + lam
            else:
                K, (offp, offl, N) = self.monolithic_implicit(Qstar, a * dt)  # :602-610
                R = np.concatenate([self.residual(i).ravel(), np.zeros(N - offp)])
                x = self.solve_pinned(K, R, offl, N)
                st["Q"] = x[:offp].reshape(nc, 2, o.nQ1)
                st["p"] = x[offp:offl].reshape(nc, o.np_)
                st["l"] = x[offl:].reshape(nf, o.nl1)
            st["p"], st["l"] = o.shift_pressure(st["p"], st["l"])  # :621
            if self.tracer is not None:  # :622-623 with _tracer_residual :415-432 (velocity of stage i throughout)
                u_i = self.tracer.project_cg(st["Q"])
                self.q[i] = self.q[0] + sum(dt * self.a_expl[i, j] * self.tracer.advection(self.q[j], u_i)
                                            for j in range(i) if self.a_expl[i, j] != 0)
        # final stage :624  (rhs in the w-row)
        Qn, _, _ = o.solve_condensed(self.final_residual(), zp, zl)
        # pressure reconstruction :629-637
        b_new = self.interp_Q(problem.f_rhs(tn + dt))
        Rp, Rl = self.reconstruction_rhs(Qn, b_new)
        _, pn, ln = o.solve_condensed(zQ, Rp, Rl)
        if self.tracer is not None:  # :638-639 with _tracer_final_residual :434-448
            self.q_tracer = self.q[0] + sum(
                dt * self.b_expl[i] * self.tracer.advection(self.q[i], self.tracer.project_cg(self.stage[i]["Q"]))
                for i in range(self.nstages) if self.b_expl[i] != 0)
        return dict(Q=Qn, p=pn, l=ln)

    def reconstruction_rhs(self, Qn, b_new):
        """hdg_imex.py:204-207:  weak_div(psi, -b + (grad Q) Q)  -  mu n.b ds"""
        o = self.o
        gphi = np.einsum("ndc,iqd->niqc", o.Jinv, o.dphiQ)
        Qq = o.eval_Q(Qn)
        gradQ = np.einsum("nci,niqd->nqcd", Qn, gphi)
        Xq = -o.eval_Q(b_new) + np.einsum("nqcd,nqd->nqc", gradQ, Qq)
        gphif = np.einsum("ndc,eiqd->neiqc", o.Jinv, o.dphiQ_f)
        Qf = o.eval_Q_facet(Qn)
        gradQf = np.einsum("nci,neiqd->neqcd", Qn, gphif)
        bf = o.eval_Q_facet(b_new)
        Xf = -bf + np.einsum("neqcd,neqd->neqc", gradQf, Qf)
        Rp = o.weak_divergence_fun(Xq, Xf)
        # - mu n.b ds on boundary facets
        Rl = np.zeros((self.mesh.nf, o.nl1))
        nb = np.einsum("nec,neqc->neq", o.normal, bf)
        flip = self.mesh.cell_flip
        for e in range(3):
            ell = o.ell[flip[:, e]]
            contrib = -o.elen[:, e, None] * np.einsum("q,nq,nmq->nm", o.wf, nb[:, e], ell)
            bnd = o.nbr[:, e] < 0
            np.add.at(Rl, self.mesh.cell_facet[bnd, e], contrib[bnd])
        return Rp, Rl

    def solve(self, problem, T_final, warmup=False, q_initial=None):
        o = self.o
        nt = 1 if warmup else int(np.round(T_final / self.dt))
        Q, p = self.initial_state(problem)  # :520-522
        lam = o.reconstruct_trace(Q, p)  # :534
        cur = dict(Q=Q, p=p, l=lam)
        if q_initial is not None:  # :523-527
            from .tracer import TracerOracle

            self.tracer = TracerOracle(o)
            self.q_tracer = self.interp_p(q_initial)
            self.q = [None] * self.nstages
        for k in range(nt):
            if self.tracer is not None:
                self.q[0] = self.q_tracer  # :559-560
            cur = self.step(cur, problem, k * self.dt)
        return cur["Q"], cur["p"]
