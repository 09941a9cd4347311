"""First-order HDG timestepper (Chorin projection or fully implicit), mirroring
`src/timesteppers/hdg_implicit.py:10-197` on the B200 engine.

Per step (reference line numbers):
  :98      Q* = project_bdm(Q)
  :100     f  = interpolate(f_rhs(t_n))
  :103-129 tentative velocity   (M + dt [adv(Q*) - flux + penalty]) Q~ = M Q + dt M f
  :133-146 mixed Poisson        a_poisson(u, phi, lambda) = -(1/dt) psi div(Q~) dx
  :150     Q <- Q~ + dt u
  :189-190 p <- phi - mean(phi)

The reference solves both systems with Firedrake's default direct LU; the engine uses BiCGStab
on the matrix-free tentative operator and the statically condensed CG trace solve, converged to
``krylov_rtol`` so that results agree with the direct solves to ~1e-10.  The Chorin right-hand side
is not in the range of the singular mixed-Poisson operator in general; the engine removes the
defect along the null vector (see oracle/hdg_oracle.py `_project_trace_rhs`).
"""

from __future__ import annotations

import tqdm

from ..auxilliary.logging import PerformanceLog
from ..auxilliary.utils import Averager
from ..functions import Function
from .common import IncompressibleEuler

__all__ = ["IncompressibleEulerHDGImplicit"]

# value at t_n of the degree-m polynomial through the m + 1 previous values (newest first)
_EXTRAPOLATION = {0: (1.0,), 1: (2.0, -1.0), 2: (3.0, -3.0, 1.0), 3: (4.0, -6.0, 4.0, -1.0),
                  4: (5.0, -10.0, 10.0, -5.0, 1.0)}


class IncompressibleEulerHDGImplicit(IncompressibleEuler):
    def __init__(self, mesh, degree, dt, flux="upwind", use_projection_method=True, callbacks=None, device=0,
                 krylov_rtol=1e-12, progress=False, preconditioner="gtmg", warm_start=False, warm_order=3):
        super().__init__(mesh, degree, dt, label="HDG Implicit", device=device, preconditioner=preconditioner)
        self.flux = flux
        assert self.flux in ["upwind", "centered"]
        self.use_projection_method = use_projection_method
        self.callbacks = [] if callbacks is None else callbacks
        self.alpha = 1  # penalty parameter (:41)
        self.engine.set_penalty(self.alpha)
        self.krylov_rtol = krylov_rtol
        # BiCGStab iterations + FGMRES iterations of the fallback (engine knobs tent_bicg_cap, tent_krylov): large
        # time steps (the reference default dt = 0.04, src/driver.py:80-86) need hundreds to thousands
        self.tentative_maxit = 5000
        # warm_start (off by default: the reference solves both systems from scratch, hdg_implicit.py:129,146):
        # both Krylov solves start from time-extrapolated guesses instead of Q^n / zero: tentative velocity
        # Q^n + sum_j c_j d^{n-j} with d = Q~ - Q and the c_j of polynomial extrapolation of degree
        # warm_order (as far as the history reaches), trace 3 lambda^{n-1} - 3 lambda^{n-2} + lambda^{n-3}.
        # The tolerances refer to the right-hand side norms, so the converged fields are the same; only the
        # iteration counts drop.
        self.warm_start = warm_start
        self.warm_order = int(warm_order)
        assert 0 <= self.warm_order <= 4
        self.progress = progress
        self.niter_tentative = Averager()
        self.niter_pressure = Averager()
        if not use_projection_method:
            from .monolithic import MonolithicStage

            self._monolithic = MonolithicStage(self)

    @PerformanceLog("tentative_velocity_solve")
    def tentative_velocity_solve(self, Q_star, rhs, Q_tentative, zero_guess):
        return self.engine.tentative_solve_dev(Q_star.data, self._dt, rhs.data, Q_tentative.data,
                                               upwind=(self.flux == "upwind"), rtol=self.krylov_rtol,
                                               zero_guess=zero_guess, maxit=self.tentative_maxit)

    @PerformanceLog("pressure_solve")
    def pressure_solve(self, Rp, u, phi, lmbda):
        return self.engine.poisson_apply_dev(None, Rp.data, None, u.data, phi.data, lmbda.data, rtol=self.krylov_rtol,
                                             maxit=2000 if self.preconditioner == "gtmg" else 100000, shift=True)

    @PerformanceLog("pressure_solve")
    def pressure_solve_update(self, Rp, Q, Q_tentative, p, lmbda):
        return self.engine.poisson_apply_update_dev(None, Rp.data, None, Q.data, p.data, lmbda.data, cq=0.0, cb=1.0,
                                                    Q_base=Q_tentative.data, cu=self._dt, cp=0.0, rtol=self.krylov_rtol,
                                                    maxit=2000 if self.preconditioner == "gtmg" else 100000)

    def initialise(self, Q_initial, p_initial):
        """interpolate the initial conditions and allocate the per-step work fields (:82-84)"""
        self.Q = self._V_Q.interpolate(Q_initial)
        self.Q.rename("velocity")
        self.p = self._V_p.interpolate(p_initial)
        self.p.rename("pressure")
        self.engine.shift_pressure_dev(self.p.data, None)  # :84
        self._Q_star, self._f, self._rhs, self._Q_tentative, self._u = (Function(self._V_Q) for _ in range(5))
        self._Rp, self._phi = Function(self._V_p), Function(self._V_p)
        self._lmbda = self._V_trace.zeros()
        self._lmbda_prev = self._V_trace.zeros()
        self._lmbda_prev2 = self._V_trace.zeros()
        # Q~ - Q of the previous steps, newest first (history of the extrapolated tentative-velocity guess)
        self._dQt_hist = [self._V_Q.zeros() for _ in range(self.warm_order + 1)] if self.warm_start else []
        self._nsteps = 0
        self.iteration_history = []  # (tentative its, pressure its) per step
        self.engine.set_initial_guess(bool(self.warm_start) and self.use_projection_method)
        self.niter_tentative.reset()
        self.niter_pressure.reset()
        return self.Q, self.p

    def initialise_tracer(self, q_initial):
        """`hdg_implicit.py:73-79`; allocates the projected velocity and the second tracer buffer"""
        q = self._tracer_initialise(q_initial)
        if q is not None:
            self._u_cg = Function(self._V_Q)
            self._q_new = Function(self._V_q, name="tracer")
        return q

    def step(self, k, f_rhs, f_field=None):
        """one timestep t_k -> t_{k+1} (:92-190).  f_field, if given, is an already interpolated
        forcing Function (host-driven forcing); otherwise f_rhs(t_k) is interpolated."""
        eng, Q, p = self.engine, self.Q, self.p
        with PerformanceLog("timestep"):
            if self.q_tracer is not None:
                # :93-96 -- `project` runs when the form is built, i.e. on the velocity of time t_k
                self._project_onto_cg(Q, self._u_cg)
            with PerformanceLog("bdm_projection"):
                self.project_bdm(Q, out=self._Q_star)  # :98
            f = f_field if f_field is not None else self._V_Q.interpolate(f_rhs(k * self._dt), out=self._f)  # :100
            eng.lincomb_dev(self._rhs.data, [(1.0, Q.data), (self._dt, f.data)])  # :126 / :182 in Riesz form
            if self.use_projection_method:
                if self.warm_start:
                    n = self._nsteps
                    # polynomial extrapolation of the increment through its last m + 1 values: binomial weights
                    m = min(self.warm_order, n - 1)
                    terms = [(1.0, Q.data)]
                    if n >= 1:
                        terms += [(_EXTRAPOLATION[m][j], self._dQt_hist[j].data) for j in range(m + 1)]
                    eng.lincomb_dev(self._Q_tentative.data, terms)
                    # the guess goes into the oldest trace buffer, which then becomes the current one
                    l1, l2, l3 = self._lmbda, self._lmbda_prev, self._lmbda_prev2
                    if n >= 3:
                        eng.lincomb_dev(l3.data, [(3.0, l1.data), (-3.0, l2.data), (1.0, l3.data)])
                    elif n == 2:
                        eng.lincomb_dev(l3.data, [(2.0, l1.data), (-1.0, l2.data)])
                    else:
                        l3.assign(l1)
                    self._lmbda, self._lmbda_prev, self._lmbda_prev2 = l3, l1, l2
                else:
                    self._Q_tentative.assign(Q)
                its = self.tentative_velocity_solve(self._Q_star, self._rhs, self._Q_tentative, zero_guess=False)  # :129
                its_t = its
                self.niter_tentative.update(its)
                if self.warm_start:  # shift the history, newest increment Q~ - Q first
                    self._dQt_hist.insert(0, self._dQt_hist.pop())
                    eng.lincomb_dev(self._dQt_hist[0].data, [(1.0, self._Q_tentative.data), (-1.0, Q.data)])
                eng.weak_divergence_dev(self._Q_tentative.data, self._Rp.data, scale=-1.0 / self._dt, mode=0)  # :145
                # :146 + :150 + :188-190 in one call: the back-substitution writes Q = Q~ + dt u and p = phi - mean
                # directly (hdg_poisson_apply_update_dev), u and phi never go to memory
                its = self.pressure_solve_update(self._Rp, Q, self._Q_tentative, p, self._lmbda)
                self.niter_pressure.update(its)
                self.iteration_history.append((its_t, its))
                self._nsteps += 1
            else:
                with PerformanceLog("unsplit_solve"):
                    self._phi.assign(p)  # initial guess: the state of the previous step
                    its = self._monolithic.solve(self._Q_star, self._dt, self._rhs, Q, self._phi, self._lmbda,
                                                 rtol=self.krylov_rtol, upwind=(self.flux == "upwind"))  # :185
                    self.niter_pressure.update(its)
            if not self.use_projection_method:
                p.assign(self._phi)  # :189-190
                eng.shift_pressure_dev(p.data, None)
            if self.q_tracer is not None:  # :192-193  q <- q + dt M^-1 adv(q, P_CG Q^k), explicit Euler
                with PerformanceLog("tracer_advection"):
                    self._tracer_advection(self.q_tracer, self._u_cg, self._q_new, c0=1.0, acc=self.q_tracer,
                                           c1=self._dt)
                self.q_tracer.data, self._q_new.data = self._q_new.data, self.q_tracer.data
        return Q, p

    def solve(self, Q_initial, p_initial, q_initial, f_rhs, T_final, warmup=False):
        nt = self.get_timesteps(T_final, warmup)
        q_tracer = self.initialise_tracer(q_initial)
        Q, p = self.initialise(Q_initial, p_initial)
        for callback in self.callbacks:
            callback.reset()
            callback(Q, p, 0, q_tracer=q_tracer)
        steps = tqdm.tqdm(range(nt)) if self.progress else range(nt)
        for k in steps:
            self.step(k, f_rhs)
            for callback in self.callbacks:
                callback(Q, p, (k + 1) * self._dt, q_tracer=q_tracer)
        return Q, p
