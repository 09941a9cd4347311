"""IMEX HDG timesteppers, mirroring `src/timesteppers/hdg_imex.py:22-1038` on the B200 engine.

The class names, constructor arguments, tableau properties and the structure of ``solve`` are the
reference's.  What changes is what sits behind ``pressure_solve`` / ``tentative_velocity_solve`` /
``project_bdm``: instead of Firedrake ``LinearVariationalSolver`` objects with
``"pc_python_type": "firedrake.SCPC"`` (`hdg_imex.py:128-170`) the statically condensed
mixed-Poisson problem is solved by the CUDA engine (K1-K5), and the tentative velocity by a
matrix-free BiCGStab.

All velocity right-hand sides are carried in Riesz form (M^-1 applied; M = detJ I in the engine's
orthonormal basis), so `_residual` (:367-391) and `_final_residual` (:393-413) become plain linear
combinations of fields:   rho_i = Q_n + sum_j a_ij/a_jj (Q_j - rho_j) + dt sum_j a^E_ij b_j.
"""

from __future__ import annotations

from abc import abstractmethod

import numpy as np
import tqdm

from ..auxilliary.logging import PerformanceLog
from ..auxilliary.utils import Averager
from ..functions import Function
from .common import IncompressibleEuler

__all__ = [
    "IncompressibleEulerHDGIMEX",
    "IncompressibleEulerHDGIMEXImplicit",
    "IncompressibleEulerHDGIMEXARS2_232",
    "IncompressibleEulerHDGIMEXARS3_443",
    "IncompressibleEulerHDGIMEXSSP2_332",
    "IncompressibleEulerHDGIMEXSSP3_433",
]


class _State:
    """(Q, p, lambda) triple, the analogue of a Function on the mixed space V (:69)"""

    def __init__(self, ts):
        self.Q = ts._V_Q.zeros()
        self.p = ts._V_p.zeros()
        self.l = ts._V_trace.zeros()

    @property
    def subfunctions(self):
        return (self.Q, self.p, self.l)

    def assign(self, other):
        self.Q.assign(other.Q)
        self.p.assign(other.p)
        self.l.assign(other.l)


class IncompressibleEulerHDGIMEX(IncompressibleEuler):
    """Abstract base class for the IMEX timesteppers (`hdg_imex.py:22-660`)"""

    def __init__(self, mesh, degree, dt, flux="upwind", use_projection_method=True, n_richardson=2, label=None,
                 callbacks=None, device=0, krylov_rtol=1e-12, tentative_rtol=None, progress=False,
                 preconditioner="gtmg"):
        super().__init__(mesh, degree, dt, label, device=device, preconditioner=preconditioner)
        self.flux = flux
        self.use_projection_method = use_projection_method
        assert self.flux in ["upwind", "centered"]
        self.alpha_penalty = 1  # :56
        self.engine.set_penalty(self.alpha_penalty)
        self.n_richardson = n_richardson
        self.callbacks = [] if callbacks is None else callbacks
        # the reference asks for rtol 1e-12 (trace, :137) and 1e-10 (tentative velocity, :226)
        self.krylov_rtol = krylov_rtol
        self.tentative_rtol = krylov_rtol if tentative_rtol is None else tentative_rtol
        self.progress = progress
        s = self.nstages
        self._stage_state = [_State(self) for _ in range(s)]  # persistent across timesteps (:72-88)
        self._Qstar = [Function(self._V_Q) for _ in range(s - 1)]
        self._Q_tentative = [self._V_Q.zeros() for _ in range(s)]
        self._b_rhs = [self._V_Q.zeros() for _ in range(s)]
        self._rho = [self._V_Q.zeros() for _ in range(s)]
        self._update = _State(self)
        self._current_state = _State(self)
        self._pressure_reconstruction = _State(self)
        self._b_new = self._V_Q.zeros()
        self._work_Q = self._V_Q.zeros()
        self._work_p = self._V_p.zeros()
        self._work_l = self._V_trace.zeros()
        self.niter_tentative = Averager()
        self.niter_pressure = Averager()
        self.niter_final_pressure = Averager()
        self.niter_pressure_reconstruction = Averager()
        if not use_projection_method:
            from .monolithic import MonolithicStage

            self._monolithic = MonolithicStage(self)

    # -- tableau (abstract) ----------------------------------------------------------------------
    @property
    @abstractmethod
    def nstages(self):
        """number of stages s"""

    @property
    @abstractmethod
    def _a_expl(self):
        """s x s explicit coefficients"""

    @property
    @abstractmethod
    def _a_impl(self):
        """s x s implicit coefficients"""

    @property
    @abstractmethod
    def _b_expl(self):
        """explicit final-stage weights"""

    @property
    @abstractmethod
    def _b_impl(self):
        """implicit final-stage weights"""

    @property
    @abstractmethod
    def _c_expl(self):
        """fractional times of the explicit term"""

    # -- helpers ---------------------------------------------------------------------------------
    def _lincomb(self, out, terms):
        """out = sum c x, chunked to the engine's 8-term kernel"""
        terms = [(c, x.data) for c, x in terms if c != 0]
        assert terms and all(x is not out.data for _, x in terms)
        self.engine.lincomb_dev(out.data, terms[:8])
        rest = terms[8:]
        while rest:
            self.engine.lincomb_dev(out.data, [(1.0, out.data)] + rest[:7])
            rest = rest[7:]

    def _compute_rho(self, i):
        """Riesz representative of `_residual(w, i)` (:367-391)"""
        assert 0 < i < self.nstages
        terms = [(1.0, self._stage_state[0].Q)]
        for j in range(1, i):
            if self._a_impl[i, j] != 0:
                c = self._a_impl[i, j] / self._a_impl[j, j]
                terms += [(c, self._stage_state[j].Q), (-c, self._rho[j])]
        for j in range(i):
            if self._a_expl[i, j] != 0:
                terms.append((self._dt * self._a_expl[i, j], self._b_rhs[j]))
        self._lincomb(self._rho[i], terms)

    def _compute_final_rho(self, out):
        """Riesz representative of `_final_residual(w)` (:393-413); note that the reference indexes
        `_b_impl[i]` for i < nstages even where the array is longer (ARS3(4,4,3), :874)"""
        terms = [(1.0, self._stage_state[0].Q)]
        for i in range(1, self.nstages):
            if self._b_impl[i] != 0:
                c = self._b_impl[i] / self._a_impl[i, i]
                terms += [(c, self._stage_state[i].Q), (-c, self._rho[i])]
        for i in range(self.nstages):
            if self._b_expl[i] != 0:
                terms.append((self._dt * self._b_expl[i], self._b_rhs[i]))
        self._lincomb(out, terms)

    @PerformanceLog("pressure_solve")
    def pressure_solve(self, key):
        """Solve the pressure-correction equation (`hdg_imex.py:257-272`).

        key is "stage_i", "final_stage" or "pressure_reconstruction"; returns the trace-solve
        iteration count (the reference reads it from condensed_ksp, :265-271)."""
        eng = self.engine
        if key.startswith("stage_"):
            i = int(key.split("_")[1])
            adt = self._a_impl[i, i] * self._dt
            eng.weak_divergence_dev(self._Q_tentative[i].data, self._work_p.data, scale=-1.0 / adt, mode=1)  # :177-179
            st = self._update
            return eng.poisson_apply_dev(None, self._work_p.data, None, st.Q.data, st.p.data, st.l.data,
                                         rtol=self.krylov_rtol, maxit=100000, shift=False)
        if key == "final_stage":
            self._compute_final_rho(self._work_Q)
            eng.mass_dev(0, self._work_Q.data, self._work_Q.data)  # dual vector (w, rho)
            st = self._current_state
            return eng.poisson_apply_dev(self._work_Q.data, None, None, st.Q.data, st.p.data, st.l.data,
                                         rtol=self.krylov_rtol, maxit=100000, shift=False)
        if key == "pressure_reconstruction":
            Q_new = self._current_state.Q
            eng.reconstruction_rhs_dev(Q_new.data, self._b_new.data, self._work_p.data, self._work_l.data)  # :204-207
            st = self._pressure_reconstruction
            return eng.poisson_apply_dev(None, self._work_p.data, self._work_l.data, st.Q.data, st.p.data, st.l.data,
                                         rtol=self.krylov_rtol, maxit=100000, shift=False)
        raise KeyError(key)

    @PerformanceLog("pressure_solve")
    def pressure_solve_update(self, i):
        """stage-i pressure correction (`hdg_imex.py:257-272`) fused with `_shift_pressure(update)` (:579) and the
        Richardson update  Q_i += Q~ + a dt u,  p_i += phi  (:580-593); the update trace stays in ``self._update.l``"""
        eng = self.engine
        adt = self._a_impl[i, i] * self._dt
        st = self._stage_state[i]
        eng.weak_divergence_dev(self._Q_tentative[i].data, self._work_p.data, scale=-1.0 / adt, mode=1)  # :177-179
        return eng.poisson_apply_update_dev(None, self._work_p.data, None, st.Q.data, st.p.data, self._update.l.data,
                                            cq=1.0, cb=1.0, Q_base=self._Q_tentative[i].data, cu=adt, cp=1.0,
                                            rtol=self.krylov_rtol, maxit=100000)

    @PerformanceLog("tentative_velocity_solve")
    def tentative_velocity_solve(self, key):
        """Compute the tentative velocity (`hdg_imex.py:274-281`, forms :232-255)"""
        eng = self.engine
        i = int(key.split("_")[1])
        adt = self._a_impl[i, i] * self._dt
        st = self._stage_state[i]
        upwind = self.flux == "upwind"
        rhs = self._work_Q
        # rhs = rho_i - Q_i + a dt M^-1 [ f_impl(w, Q_i; Q*) + g(w, p_i, lambda_i) ]   (:239-247)
        eng.fimpl_apply_dev(self._Qstar[i - 1].data, st.Q.data, rhs.data, c0=-1.0, c1=adt, upwind=upwind)
        eng.pressure_gradient_dev(st.p.data, st.l.data, rhs.data, c0=1.0, c1=adt)
        eng.lincomb_dev(rhs.data, [(1.0, rhs.data), (1.0, self._rho[i].data)])
        return eng.tentative_solve_dev(self._Qstar[i - 1].data, adt, rhs.data, self._Q_tentative[i].data, upwind=upwind,
                                       rtol=self.tentative_rtol, zero_guess=True)

    def _shift_pressure(self, state):
        """`_shift_pressure` (:471-478)"""
        self.engine.shift_pressure_dev(state.p.data, state.l.data)

    def _reconstruct_trace(self, state):
        """`_reconstruct_trace` (:450-469)"""
        self.engine.reconstruct_trace_dev(state.Q.data, state.p.data, state.l.data)

    def solve(self, Q_initial, p_initial, q_initial, f_rhs, T_final, warmup=False):
        eng = self.engine
        nt = self.get_timesteps(T_final, warmup)
        q_tracer = self._tracer_initialise(q_initial)  # :523-529
        if q_tracer is not None:
            self._q = [Function(self._V_q) for _ in range(self.nstages)]  # :81-85
            self._u_cg = [Function(self._V_Q) for _ in range(self.nstages)]
        cur = self._current_state
        self._V_Q.interpolate(Q_initial, out=cur.Q)  # :520
        self._V_p.interpolate(p_initial, out=cur.p)  # :521
        eng.shift_pressure_dev(cur.p.data, None)  # :522
        cur.Q.rename("Q")
        cur.p.rename("p")
        self._reconstruct_trace(cur)  # :534
        for a in (self.niter_tentative, self.niter_pressure, self.niter_final_pressure,
                  self.niter_pressure_reconstruction):
            a.reset()
        for callback in self.callbacks:
            callback.reset()
            callback(cur.Q, cur.p, 0, q_tracer=q_tracer)
        steps = tqdm.tqdm(range(nt)) if self.progress else range(nt)
        for k in steps:
            with PerformanceLog("timestep"):
                tn = k * self._dt
                for i in range(self.nstages):  # :554-557
                    self._V_Q.interpolate(f_rhs(tn + self._c_expl[i] * self._dt), out=self._b_rhs[i])
                self._stage_state[0].assign(cur)  # :558
                if q_tracer is not None:
                    self._q[0].assign(q_tracer)  # :559-560
                for i in range(1, self.nstages):
                    st = self._stage_state[i]
                    adt = self._a_impl[i, i] * self._dt
                    with PerformanceLog("bdm_projection"):  # :564-567
                        self.project_bdm(self._stage_state[i - 1].Q, out=self._Qstar[i - 1])
                    self._compute_rho(i)
                    if self.use_projection_method:
                        for _ in range(self.n_richardson):  # :570
                            its = self.tentative_velocity_solve(f"stage_{i:d}")
                            self.niter_tentative.update(its)
                            # :575-593 in one call: pressure solve, _shift_pressure(update) (:579) and the velocity /
                            # pressure updates (:580-593) inside the back-substitution kernel
                            its = self.pressure_solve_update(i)
                            self.niter_pressure.update(its)
                            eng.lincomb_dev(st.l.data, [(1.0, st.l.data), (1.0, self._update.l.data)])  # :594-599
                    else:
                        with PerformanceLog("unsplit_solve"):  # :601-620
                            self._monolithic.solve(self._Qstar[i - 1], adt, self._rho[i], st.Q, st.p, st.l,
                                                   rtol=self.krylov_rtol, upwind=(self.flux == "upwind"))
                    self._shift_pressure(st)  # :621
                    if q_tracer is not None:  # :622-623
                        self._tracer_stage(i)
                its = self.pressure_solve("final_stage")  # :624
                self.niter_final_pressure.update(its)
                self._V_Q.interpolate(f_rhs(tn + self._dt), out=self._b_new)  # :629
                its = self.pressure_solve("pressure_reconstruction")  # :630
                self.niter_pressure_reconstruction.update(its)
                cur.p.assign(self._pressure_reconstruction.p)  # :633-636
                cur.l.assign(self._pressure_reconstruction.l)
                self._shift_pressure(cur)  # :637
                if q_tracer is not None:  # :638-639
                    self._tracer_final(q_tracer)
            for callback in self.callbacks:
                callback(cur.Q, cur.p, tn + self._dt, q_tracer=q_tracer)
        self.print_iteration_summary()
        return cur.Q, cur.p

    def _tracer_accumulate(self, out, terms):
        """out = q_0 + sum_t c_t M^-1 adv(q_t, u_t) for terms = [(c, q, u_cg), ...]"""
        if not terms:
            return out.assign(self._q[0])
        acc = self._q[0]
        for c, q, u in terms:
            self._tracer_advection(q, u, out, c0=1.0, acc=acc, c1=c)
            acc = out
        return out

    @PerformanceLog("tracer_advection")
    def _tracer_stage(self, i):
        """`solve(a_tracer == self._tracer_residual(chi, i), self._q[i])` (`hdg_imex.py:415-432,622-623`):
        q_i = q_0 + dt sum_{j<i} a^E_ij M^-1 adv(q_j, P_CG Q_i) -- every term uses the velocity of
        stage i, as the reference does"""
        self._project_onto_cg(self._stage_state[i].Q, self._u_cg[i])
        terms = [(self._dt * self._a_expl[i, j], self._q[j], self._u_cg[i]) for j in range(i)
                 if self._a_expl[i, j] != 0]
        self._tracer_accumulate(self._q[i], terms)

    @PerformanceLog("tracer_advection")
    def _tracer_final(self, q_tracer):
        """`solve(a_tracer == self._tracer_final_residual(chi), q_tracer)` (`hdg_imex.py:434-448,638-639`):
        q^{n+1} = q_0 + dt sum_i b^E_i M^-1 adv(q_i, P_CG Q_i)"""
        if self._b_expl[0] != 0:  # stage 0 is the state at t_n; stages >= 1 were projected in _tracer_stage
            self._project_onto_cg(self._stage_state[0].Q, self._u_cg[0])
        terms = [(self._dt * self._b_expl[i], self._q[i], self._u_cg[i]) for i in range(self.nstages)
                 if self._b_expl[i] != 0]
        self._tracer_accumulate(q_tracer, terms)

    def print_iteration_summary(self, file=None):
        """:648-659"""
        print("average number of solver iterations", file=file)
        print(40 * "-", file=file)
        print(f"  tentative velocity its      : {self.niter_tentative.value:8.2f}", file=file)
        if self.use_projection_method:
            print(f"  pressure its                : {self.niter_pressure.value:8.2f}", file=file)
            print(f"  final pressure its          : {self.niter_final_pressure.value:8.2f}", file=file)
        print(f"  pressure reconstruction its : {self.niter_pressure_reconstruction.value:8.2f}", file=file)
        print(file=file)


#######################################################################################
#       S P E C I F I C     I M E X     T I M E S T E P P E R S                       #
#######################################################################################


def _make(label, nstages, a_expl, a_impl, b_expl, b_impl, c_expl, doc):
    """build a tableau subclass with the reference's constructor signature (:671-700 etc.)"""

    class _IMEX(IncompressibleEulerHDGIMEX):
        def __init__(self, mesh, degree, dt, flux="upwind", use_projection_method=True, n_richardson=2,
                     callbacks=None, **engine_options):
            super().__init__(mesh, degree, dt, flux, use_projection_method, n_richardson, label=label,
                             callbacks=callbacks, **engine_options)

        nstages = property(lambda self: nstages)
        _a_expl = property(lambda self: np.asarray(a_expl, dtype=float))
        _a_impl = property(lambda self: np.asarray(a_impl, dtype=float))
        _b_expl = property(lambda self: np.asarray(b_expl, dtype=float))
        _b_impl = property(lambda self: np.asarray(b_impl, dtype=float))
        _c_expl = property(lambda self: np.asarray(c_expl, dtype=float))

    _IMEX.__doc__ = doc
    return _IMEX


_g = 1 - 1 / np.sqrt(2)
_d = -2 / 3 * np.sqrt(2)
_al, _be, _et = 0.24169426078821, 0.06042356519705, 0.12915286960590
_de = 1 / 2 - _al - _be - _et

#: first order implicit method in IMEX form (:668-729)
IncompressibleEulerHDGIMEXImplicit = _make(
    "HDG IMEX Implicit", 2, [[0, 0], [1, 0]], [[0, 0], [0, 1]], [1, 0], [0, 1], [0, 1],
    "IMEX implementation of the first order implicit method")
IncompressibleEulerHDGIMEXImplicit.__name__ = "IncompressibleEulerHDGIMEXImplicit"

#: ARS2(2,3,2) (:732-799)
IncompressibleEulerHDGIMEXARS2_232 = _make(
    "HDG IMEX ARS2(2,3,2)", 3, [[0, 0, 0], [_g, 0, 0], [_d, 1 - _d, 0]], [[0, 0, 0], [0, _g, 0], [0, 1 - _g, _g]],
    [0, 1 - _g, _g], [0, 1 - _g, _g], [0, _g, 1], "IMEX ARS2(2,3,2) timestepper")
IncompressibleEulerHDGIMEXARS2_232.__name__ = "IncompressibleEulerHDGIMEXARS2_232"

#: ARS3(4,4,3) (:802-879); _b_impl keeps the reference's six entries
IncompressibleEulerHDGIMEXARS3_443 = _make(
    "HDG IMEX ARS3(4,4,3)", 5,
    [[0, 0, 0, 0, 0], [1 / 2, 0, 0, 0, 0], [11 / 18, 1 / 18, 0, 0, 0], [5 / 6, -5 / 6, 1 / 2, 0, 0],
     [1 / 4, 7 / 4, 3 / 4, -7 / 4, 0]],
    [[0, 0, 0, 0, 0], [0, 1 / 2, 0, 0, 0], [0, 1 / 6, 1 / 2, 0, 0], [0, -1 / 2, 1 / 2, 1 / 2, 0],
     [0, 3 / 2, -3 / 2, 1 / 2, 1 / 2]],
    [1 / 4, 7 / 4, 3 / 4, -7 / 4, 0], [0, 3 / 2, -3, 2, 1 / 2, 1 / 2], [0, 1 / 2, 2 / 3, 1 / 2, 1],
    "IMEX ARS3(4,4,3) timestepper")
IncompressibleEulerHDGIMEXARS3_443.__name__ = "IncompressibleEulerHDGIMEXARS3_443"

#: SSP2(3,3,2) (:882-949), the driver default (driver.py:134)
IncompressibleEulerHDGIMEXSSP2_332 = _make(
    "HDG IMEX SSP2(3,3,2)", 3, [[0, 0, 0], [1 / 2, 0, 0], [1 / 2, 1 / 2, 0]],
    [[1 / 4, 0, 0], [0, 1 / 4, 0], [1 / 3, 1 / 3, 1 / 3]], [1 / 3, 1 / 3, 1 / 3], [1 / 3, 1 / 3, 1 / 3], [0, 1, 1 / 2],
    "IMEX SSP2(3,3,2) timestepper")
IncompressibleEulerHDGIMEXSSP2_332.__name__ = "IncompressibleEulerHDGIMEXSSP2_332"

#: SSP3(4,3,3) (:952-1038), coefficients of Pareschi & Russo (2005)
IncompressibleEulerHDGIMEXSSP3_433 = _make(
    "HDG IMEX SSP3(4,3,3)", 4, [[0, 0, 0, 0], [0, 0, 0, 0], [0, 1, 0, 0], [0, 1 / 4, 1 / 4, 0]],
    [[_al, 0, 0, 0], [-_al, _al, 0, 0], [0, 1 - _al, _al, 0], [_be, _et, _de, _al]], [0, 1 / 6, 1 / 6, 2 / 3],
    [0, 1 / 6, 1 / 6, 2 / 3], [0, 0, 1, 1 / 2], "IMEX SSP3(4,3,3) timestepper")
IncompressibleEulerHDGIMEXSSP3_433.__name__ = "IncompressibleEulerHDGIMEXSSP3_433"
