"""ctypes front end of oracle/cpu_ref/hdg_cpu_ref.cpp  --  TEST / MEASUREMENT INFRASTRUCTURE ONLY.

The compiled (C++/OpenMP) restatement of the reference's Chorin step that bench.py times as ``cpu_baseline`` and as
``--impl reference``.  Only tests/, bench.py's CPU legs and ``__graft_entry__.build()`` import this
module; the product never does.  The reference-element tabulation is the numpy oracle's (HDGOracle._tabulate), so the
fields of the two checkers are directly comparable (tests/test_cpu_ref.py, 1e-10).
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_ref", "hdg_cpu_ref.cpp")
LIB = os.path.join(HERE, "cpu_ref", "libhdg_cpu_ref.so")

_c_dp = ctypes.POINTER(ctypes.c_double)
_c_ip = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> str:
    """g++ -O3 -fopenmp -> oracle/cpu_ref/libhdg_cpu_ref.so (AVX2/FMA code, runs on any x86-64 box of this decade)"""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = ["g++", "-O3", "-mavx2", "-mfma", "-fopenmp", "-std=c++17", "-shared", "-fPIC", "-o", tmp, SRC]
    subprocess.run(cmd, check=True)
    os.replace(tmp, LIB)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        lib.hdgcpu_create.restype = ctypes.c_void_p
        lib.hdgcpu_create.argtypes = ([ctypes.c_int] * 3 + [_c_dp] + [_c_ip] * 4 + [ctypes.c_int] * 3 + [_c_dp] * 11
                                      + [ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int])
        lib.hdgcpu_destroy.argtypes = [ctypes.c_void_p]
        lib.hdgcpu_set_tentative_precond.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        lib.hdgcpu_threads.restype = ctypes.c_int
        lib.hdgcpu_precond_apply.argtypes = [ctypes.c_void_p, ctypes.c_double, _c_dp, _c_dp]
        lib.hdgcpu_project_bdm.argtypes = [ctypes.c_void_p, _c_dp, _c_dp]
        lib.hdgcpu_fimpl_apply.argtypes = [ctypes.c_void_p, _c_dp, _c_dp, _c_dp]
        lib.hdgcpu_local_schur.argtypes = [ctypes.c_void_p, _c_dp]
        lib.hdgcpu_tentative_solve.argtypes = [ctypes.c_void_p, _c_dp, ctypes.c_double, _c_dp, _c_dp, ctypes.c_double,
                                               ctypes.c_int, _c_ip]
        lib.hdgcpu_poisson_solve.argtypes = [ctypes.c_void_p, _c_dp, _c_dp, _c_dp, ctypes.c_double, ctypes.c_int, _c_dp,
                                             _c_dp, _c_dp, _c_ip]
        lib.hdgcpu_chorin_step.argtypes = [ctypes.c_void_p, _c_dp, _c_dp, _c_dp, ctypes.c_double, ctypes.c_double,
                                           ctypes.c_int, _c_ip]
        lib.hdgcpu_timers.argtypes = [ctypes.c_void_p, _c_dp, ctypes.POINTER(ctypes.c_int64)]
        _lib = lib
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(_c_dp)


def _ip(a):
    return a.ctypes.data_as(_c_ip)


def _f8(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class ChorinCpuRef:
    """`IncompressibleEulerHDGImplicit(use_projection_method=True)` (hdg_implicit.py:92-190) on the host cores"""

    LABELS = ("timestep", "bdm_projection", "tentative_velocity_solve", "pressure_solve")

    def __init__(self, mesh, degree, dt, flux="upwind", tau=1.0, alpha=1.0, rtol=1e-12, nthreads=0, maxit=2000,
                 precond=1, sweeps=20):
        from .hdg_oracle import HDGOracle

        self.lib = load()
        self.mesh, self.k, self.dt, self.rtol, self.maxit = mesh, degree, dt, rtol, maxit
        # tabulation only: geometry and all operators are rebuilt in C++
        o = HDGOracle.__new__(HDGOracle)
        o.mesh, o.k, o.tau, o.alpha, o.flux = mesh, degree, float(tau), float(alpha), flux
        from incompressibleeulerhdg_b200 import refelem as R

        o.nQ1, o.np_, o.nl1 = R.ncell(degree + 1), R.ncell(degree), degree + 1
        o.nQ, o.nl = 2 * o.nQ1, 3 * (degree + 1)
        o.nA = o.nQ + o.np_
        o.nq_facet = (3 * degree + 4 + 1) // 2
        o._tabulate()
        HDGOracle._bdm_setup(o)
        self.tab = o
        self.nQ1, self.np_, self.nl1 = o.nQ1, o.np_, o.nl1
        self._keep = [_f8(mesh.cell_xy), np.ascontiguousarray(mesh.cell_facet, dtype=np.int32),
                      np.ascontiguousarray(mesh.cell_flip, dtype=np.int32),
                      np.ascontiguousarray(mesh.facet_cell, dtype=np.int32),
                      np.ascontiguousarray(mesh.facet_local, dtype=np.int32)]
        tabs = [_f8(o.wq), _f8(o.phiQ), _f8(o.dphiQ), _f8(o.phiP), _f8(o.wf), _f8(o.phiQ_f), _f8(o.phiP_f), _f8(o.ell),
                _f8(o._bdm_facet), _f8(o._bdm_int),
                _f8(np.array([R.legendre01(degree + 1, o.sq), R.legendre01(degree + 1, 1.0 - o.sq)]))]
        xy, cf, cfl, fc, fl = self._keep
        self.h = self.lib.hdgcpu_create(degree, mesh.nc, mesh.nf, _dp(xy), _ip(cf), _ip(cfl), _ip(fc), _ip(fl),
                                        len(o.wq), len(o.wf), o._bdm_int.shape[0], *[_dp(t) for t in tabs], float(tau),
                                        float(alpha), 1 if flux == "upwind" else 0, int(nthreads))
        if not self.h:
            raise RuntimeError("hdgcpu_create failed (singular local operator or unsupported degree)")
        self.threads = self.lib.hdgcpu_threads()
        self.lib.hdgcpu_set_tentative_precond(self.h, int(precond), int(sweeps))
        self.iterations = []

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.hdgcpu_destroy(self.h)
            self.h = None

    # -- stages (for the cross-checks against the numpy oracle) -------------------------------------------------------
    def project_bdm(self, Q):
        Q = _f8(Q)
        out = np.empty_like(Q)
        assert self.lib.hdgcpu_project_bdm(self.h, _dp(Q), _dp(out)) == 0
        return out

    def f_impl_apply(self, Q, Qstar):
        Q, Qstar = _f8(Q), _f8(Qstar)
        out = np.empty_like(Q)
        self.lib.hdgcpu_fimpl_apply(self.h, _dp(Q), _dp(Qstar), _dp(out))
        return out

    def precond_apply(self, adt, r):
        r = _f8(r)
        out = np.empty_like(r)
        assert self.lib.hdgcpu_precond_apply(self.h, float(adt), _dp(r), _dp(out)) == 0
        return out

    def local_schur(self):
        nl = 3 * self.nl1
        out = np.empty((self.mesh.nc, nl, nl))
        self.lib.hdgcpu_local_schur(self.h, _dp(out))
        return out

    def tentative_solve(self, Qstar, adt, rhs, x0=None):
        Qstar, rhs = _f8(Qstar), _f8(rhs)
        x = np.zeros_like(rhs) if x0 is None else _f8(x0).copy()
        it = ctypes.c_int32(0)
        rc = self.lib.hdgcpu_tentative_solve(self.h, _dp(Qstar), float(adt), _dp(rhs), _dp(x), self.rtol, self.maxit,
                                             ctypes.byref(it))
        if rc:
            raise RuntimeError(f"tentative BiCGStab did not converge (rc={rc}, {it.value} iterations)")
        return x, it.value

    def poisson_solve(self, Ru, Rp, Rl):
        nc, nf = self.mesh.nc, self.mesh.nf
        Ru = None if Ru is None else _f8(Ru)
        Rp = None if Rp is None else _f8(Rp)
        Rl = None if Rl is None else _f8(Rl)
        u, p, lam = np.empty((nc, 2, self.nQ1)), np.empty((nc, self.np_)), np.empty((nf, self.nl1))
        it = ctypes.c_int32(0)
        rc = self.lib.hdgcpu_poisson_solve(self.h, _dp(Ru), _dp(Rp), _dp(Rl), self.rtol, 100 * self.maxit, _dp(u), _dp(p),
                                           _dp(lam), ctypes.byref(it))
        if rc:
            raise RuntimeError(f"trace CG did not converge ({it.value} iterations)")
        return u, p, lam, it.value

    # -- time stepping ---------------------------------------------------------------------------------------------------
    def initial_state(self, problem):
        from .hdg_oracle import HDGOracle

        o = HDGOracle(self.mesh, self.k)
        Q = o.interpolate_cell(problem.Q_stationary, "Q")
        p = o.interpolate_cell(problem.p_stationary, "p")
        p = p - o.integral_p(p) / self.mesh.volume * o.const_p()
        self._interp = o
        return _f8(Q), _f8(p)

    def forcing(self, f_fun):
        return _f8(self._interp.interpolate_cell(f_fun, "Q"))

    def step(self, Q, p, f):
        """one timestep in place; `f` = interpolated forcing coefficients [nc,2,nQ1]"""
        its = (ctypes.c_int32 * 2)(0, 0)
        rc = self.lib.hdgcpu_chorin_step(self.h, _dp(Q), _dp(p), _dp(f), float(self.dt), self.rtol, self.maxit, its)
        if rc:
            raise RuntimeError(f"hdgcpu_chorin_step failed (rc={rc}, iterations {its[0]}, {its[1]})")
        self.iterations.append((its[0], its[1]))
        return Q, p

    def solve(self, problem, T_final):
        nt = int(np.round(T_final / self.dt))
        Q, p = self.initial_state(problem)
        for k in range(nt):
            self.step(Q, p, self.forcing(problem.f_rhs(k * self.dt)))
        return Q, p

    def timers(self):
        sec = (ctypes.c_double * 4)()
        n = (ctypes.c_int64 * 4)()
        self.lib.hdgcpu_timers(self.h, sec, n)
        return {lab: {"seconds": sec[i], "calls": n[i]} for i, lab in enumerate(self.LABELS)}
