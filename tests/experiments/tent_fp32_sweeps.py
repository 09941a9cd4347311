"""Development tool (not collected by pytest): FP32 storage of the Chebyshev sweep vectors of the tentative-velocity
preconditioner, on the host-compiled device kernels of tests/test_tent_host.py.

Result (k = 2, nx = 6, upwind, rtol 1e-12; iterations / error against the oracle's sparse-direct solve):

                                    present update  x = Phat^-1 y        flexible update  x += alpha xhat(p) + omega xhat(s)
    FP64 sweeps, cell blocks 0/1    85 / 2e-13,  42 / 2e-13              85 / 2e-13,  42 / 2e-13
    FP32 sweeps, cell blocks 0/1    96 / 2e-08,  42 / 7e-09              96 / 3e-13,  42 / 2e-13

i.e. rounding the sweep vectors to FP32 does not cost iterations, but with the present recovery of x from the
accumulated y the attainable accuracy drops to ~1e-8 (the preconditioner is no longer exactly linear); accumulating x
from the preconditioned directions restores round-off accuracy.  So FP32 sweeps need the flexible update first.

    python tests/experiments/tent_fp32_sweeps.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import test_tent_host as T
import scipy.sparse as sp, scipy.sparse.linalg as spla
lib = T.host_build.build("tent_host.cpp", tempfile.mkdtemp())
mesh, o, Q0, Qs, adt = T._problem(2, 6, "upwind")
b = Q0 + 0.01 * np.random.default_rng(11).standard_normal(Q0.shape)
M = sp.diags(np.repeat(o.detJ, o.nQ))
x_ref = spla.spsolve((M - adt * o.f_impl_matrix(Qs)).tocsc(), M @ b.ravel()).reshape(b.shape)
class FP32Sweeps(T.HostTentative):
    def sweep(self, inv_aalpha, rhs, x, cd, cr, zero, mode):
        if mode == 0:   # preconditioner sweeps: vectors stored in FP32 (rounded on every store)
            rhs = None if rhs is None else rhs.astype(np.float32).astype(np.float64)
            x = x.astype(np.float32).astype(np.float64)
            self.d[:] = self.d.astype(np.float32).astype(np.float64)
            out = super().sweep(inv_aalpha, rhs, x, cd, cr, zero, mode)
            self.d[:] = self.d.astype(np.float32).astype(np.float64)
            return out.astype(np.float32).astype(np.float64)
        return super().sweep(inv_aalpha, rhs, x, cd, cr, zero, mode)
for cls in (T.HostTentative, FP32Sweeps):
    for cb in (False, True):
        ht = cls(lib, mesh, 2)
        x, its = ht.solve(T.soa(Qs), adt, True, T.soa(b), 1e-12, cb)
        err = np.abs(T.aos(x, o.nQ1) - x_ref).max() / np.abs(x_ref).max()
        print(cls.__name__, "cell blocks", int(cb), "iterations", its, "error vs direct solve %.1e" % err, flush=True)

def solve_flexible(ht, Qstar, adt, upwind, b, rtol, cb, maxit=400):
    """BiCGStab with the solution updated by the preconditioned directions: x += alpha xhat(p) + omega xhat(s)"""
    nq, nmu = 2 * ht.nq1 * ht.nc, ht.nm * ht.nf
    inv_aalpha = 1.0 / (adt * ht.alpha)
    ht.cellblock = None
    if cb:
        work = np.zeros((ht.nq1 * ht.nq1, ht.nc)); ht.cellblock = np.zeros((ht.nq1 * ht.nq1, ht.nc), np.float32)
        ht.lib.th_advblock(ht.k, int(upwind), ht.nc, T.dp(ht.xy), T.ip(ht.nbr), T.dp(Qstar), T.ctypes.c_double(adt), T.dp(work), ht.cellblock.ctypes.data_as(T.FP))
    def split(v):
        return (np.ascontiguousarray(v[:nq].reshape(2 * ht.nq1, ht.nc)), np.ascontiguousarray(v[nq:].reshape(ht.nm, ht.nf)))
    def op(v):
        vx, vmu = split(v)
        in_x = ht.scaled_x(vx)
        mu, nyx = ht.precond_x(inv_aalpha, in_x, vmu)
        xh = ht.xhat(in_x, mu)
        out_x = ht.fimpl(upwind, Qstar, xh, 1.0, -adt, Z=in_x, alpha=0.0)
        out_mu = ht.sweep(inv_aalpha, nyx, mu, 0.0, 0.0, 0, 1)
        return np.concatenate([out_x.ravel(), out_mu.ravel()]), xh
    r = np.concatenate([b.ravel(), np.zeros(nmu)])
    bb = float(b.ravel() @ b.ravel())
    rhat, p = r.copy(), r.copy()
    x = np.zeros_like(b)
    rho, its = float(rhat @ r), 0
    while float(r @ r) > rtol * rtol * bb and its < maxit:
        its += 1
        v, xh_p = op(p)
        al = rho / float(rhat @ v)
        s = r - al * v
        t, xh_s = op(s)
        om = float(t @ s) / float(t @ t)
        x += al * xh_p + om * xh_s
        r = s - om * t
        rho_new = float(rhat @ r)
        p = r + (rho_new / rho) * (al / om) * (p - om * v)
        rho = rho_new
    return x, its

print("flexible update (x accumulated from the preconditioned directions):")
for cls in (T.HostTentative, FP32Sweeps):
    for cb in (False, True):
        ht = cls(lib, mesh, 2)
        x, its = solve_flexible(ht, T.soa(Qs), adt, True, T.soa(b), 1e-12, cb)
        err = np.abs(T.aos(x, o.nQ1) - x_ref).max() / np.abs(x_ref).max()
        print(cls.__name__, "cell blocks", int(cb), "iterations", its, "error vs direct solve %.1e" % err, flush=True)
