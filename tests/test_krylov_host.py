"""CPU execution of the vector / reduction kernels of the two Krylov solvers (csrc/hdg_krylov.cuh: trace CG with the
blocked-ELL SpMV `k_cg_spmv` -- the roofline kernel of bench.py --, null-space handling, BiCGStab; SURVEY.md §8 a4,
a7-a9), compiled with g++ through tests/host_kernels (test infrastructure; the engine has no CPU path) and run with one
block of one thread.  The launch sequences of `run_pcg_mg` and `bicgstab_loop` (csrc/hdg_engine.cu) are replayed
kernel by kernel, the Krylov scalars stay in the library as they stay in device memory in the engine.  With this file
every arithmetic operation of a Chorin step is covered on the CPU by the engine's own device code."""
import ctypes
import os
import sys

import numpy as np
import pytest

from incompressibleeulerhdg_b200.mesh import UnitSquareMesh
from oracle.hdg_oracle import HDGOracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_kernels"))
import build as host_build  # noqa: E402
from test_mg_host import HostGTMG  # noqa: E402
from test_poisson_host import HostMesh, dp, ip  # noqa: E402
from test_tent_host import HostTentative, _problem, aos, soa  # noqa: E402

cd, sz = ctypes.c_double, ctypes.c_size_t


@pytest.fixture(scope="module")
def libs(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host_kernels"))
    return {n: host_build.build(n + "_host.cpp", out) for n in ("poisson", "mg", "tent", "krylov")}


def cg_state(lk):
    rz0, rz, it, done = cd(0), cd(0), ctypes.c_int(0), ctypes.c_int(0)
    lk.kh_cg_state(ctypes.byref(rz0), ctypes.byref(rz), ctypes.byref(it), ctypes.byref(done))
    return rz0.value, rz.value, it.value, done.value


def pcg_mg_kernels(lk, mg, bvec, mean_part, rtol, maxit=200):
    """run_pcg_mg (zero guess) with the engine's CG kernels; `mg` supplies the ELL matrix and the V-cycle"""
    b, nf = mg.b, mg.nf
    r = np.ascontiguousarray(bvec.copy())
    x, z, p, q = (np.zeros((b, nf)) for _ in range(4))
    part_mean, part_pq, part_rz, part_q0 = (np.zeros(1) for _ in range(4))
    part_mean[0] = mean_part
    assert lk.kh_cg_init(b, nf, dp(mg.dinv), dp(part_mean), dp(r), dp(x), dp(z), dp(p), dp(part_rz)) == 0
    z = np.ascontiguousarray(mg.apply(r))
    lk.kh_dot2(sz(b * nf), dp(r), dp(z), None, dp(part_rz), None)
    p[:] = z
    lk.kh_mode0_partial(nf, dp(z), dp(part_mean))
    lk.kh_sub_mode0(nf, dp(part_mean), dp(p))
    lk.kh_cg_start(dp(part_rz), cd(rtol), maxit)
    while True:
        _, _, it, done = cg_state(lk)
        if done or it >= maxit:
            break
        assert lk.kh_cg_spmv(b, nf, dp(mg.val), ip(mg.col), dp(p), dp(q), dp(part_pq), 1) == 0
        lk.kh_mode0_partial(nf, dp(q), dp(part_q0))
        lk.kh_cg_update_plain(b, nf, dp(p), dp(q), dp(x), dp(r), dp(part_pq), dp(part_q0))
        z = np.ascontiguousarray(mg.apply(r))
        lk.kh_dot2(sz(b * nf), dp(r), dp(z), None, dp(part_rz), None)
        lk.kh_mode0_partial(nf, dp(z), dp(part_mean))
        assert lk.kh_cg_pupdate(b, nf, dp(z), dp(p), dp(part_rz), dp(part_mean)) == 0
    return x, cg_state(lk)


@pytest.mark.parametrize("k,nx", [(1, 8), (2, 6), (3, 4)])
def test_trace_cg_kernels_on_the_host(libs, k, nx):
    mesh = UnitSquareMesh(nx, perturb=0.1)
    mg = HostGTMG((libs["poisson"], libs["mg"]), mesh, k)
    lk, hm, o = libs["krylov"], HostMesh(mesh), HDGOracle(mesh, k)
    b, nf, nc = mg.b, mesh.nf, mesh.nc
    rng = np.random.default_rng(8)
    # the SpMV against the oracle's assembled trace matrix
    x = rng.standard_normal((b, nf))
    q, part = np.zeros((b, nf)), np.zeros(1)
    assert lk.kh_cg_spmv(b, nf, dp(mg.val), ip(mg.col), dp(x), dp(q), dp(part), 0) == 0
    S = o.assemble_trace_matrix()
    ref = -(S @ x.T.ravel()).reshape(nf, b).T
    assert np.abs(q - ref).max() < 1e-11 * np.abs(ref).max()
    assert abs(part[0] - np.sum(x * q)) < 1e-10 * abs(np.sum(x * q))          # fused <p, P p>
    # forward elimination -> k_trace_rhs -> CG kernels + V-cycle -> compare with the oracle's condensed solve
    Rp = rng.standard_normal((nc, o.np_))
    gK = np.zeros((3 * b, nc))
    assert libs["poisson"].ph_forward(k, nc, dp(hm.xy), ip(hm.cell_flip), cd(1.0), None,
                                      dp(np.ascontiguousarray(Rp.T)), dp(gK)) == 0
    bvec, pm = np.zeros((b, nf)), np.zeros(1)
    assert lk.kh_trace_rhs(k, nc, nf, dp(gK), None, ip(hm.facet_cell), ip(hm.facet_local), dp(bvec), dp(pm)) == 0
    assert abs(pm[0] - bvec[0].sum()) < 1e-10 * max(1.0, np.abs(bvec[0]).sum())
    _, _, _, parts = o.solve_condensed(np.zeros((nc, 2, o.nQ1)), Rp, np.zeros((nf, b)), return_parts=True)
    lam, (rz0, rz, its, done) = pcg_mg_kernels(lk, mg, bvec, pm[0], 1e-12)
    assert done == 1 and 0 < its <= 30 and rz <= 1e-24 * rz0
    res = S @ lam.T.ravel() - parts["r"].ravel()
    assert np.linalg.norm(res) < 1e-9 * np.linalg.norm(parts["r"])
    lam_np, its_np = mg.pcg(bvec, mg.apply, rtol=1e-12)                     # the numpy CG of test_mg_host.py
    d = lam - lam_np
    d[0] -= d[0].mean()
    assert np.abs(d).max() < 1e-9 * np.abs(lam_np).max() and abs(its - its_np) <= 2
    # _shift_pressure: k_pmean_partial + k_shift against the oracle
    p = rng.standard_normal((nc, o.np_))
    lam2 = rng.standard_normal((nf, b))
    ps, ls = np.ascontiguousarray(p.T), np.ascontiguousarray(lam2.T)
    assert lk.kh_shift_pressure(nc, nf, dp(hm.xy), cd(mesh.volume), dp(ps), dp(ls), dp(np.zeros(1))) == 0
    po, lo = o.shift_pressure(p, lam2)
    assert np.abs(ps.T - po).max() < 1e-12 and np.abs(ls.T - lo).max() < 1e-12
    # layout conversion used by hdg_upload / hdg_download
    a = rng.standard_normal((nc, 7))
    s_, back = np.zeros((7, nc)), np.zeros((nc, 7))
    lk.kh_aos_to_soa(dp(a), dp(s_), nc, 7)
    lk.kh_soa_to_aos(dp(s_), dp(back), nc, 7)
    assert np.array_equal(s_, a.T) and np.array_equal(back, a)


def bicgstab_kernels(lk, op, n, r0, bb, rtol, maxit, flex=None, resume=()):
    """bicgstab_loop with the engine's BiCGStab kernels; returns the accumulated update y and the iteration count.
    flex = (xh, x, nx) selects the flexible variant ("tent_flex"): `op` leaves [Phat^-1 in]_x in the array xh and
    the solution x (velocity part, nx entries) is accumulated from these directions instead of y."""
    r = np.ascontiguousarray(r0.copy())
    rhat, p, v, sv, t, y = (np.zeros(n) for _ in range(6))
    p_rv, p_ts, p_tt, p_rho, p_rr, p_bb = (np.zeros(1) for _ in range(6))
    p_bb[0] = bb
    lk.kh_bi_init(sz(n), dp(r), dp(r), dp(rhat), dp(p), dp(p_rr))
    lk.kh_bi_start(dp(p_rr), dp(p_bb), cd(rtol), maxit)
    it, done = ctypes.c_int(0), ctypes.c_int(0)
    resume = list(resume)  # tolerance factors applied, one after the other, whenever the run reports convergence
    while True:
        lk.kh_bi_state(ctypes.byref(it), ctypes.byref(done))
        if done.value == 1 and resume and it.value < maxit:
            lk.kh_bi_resume(sz(n), dp(r), dp(v), dp(p), dp(p_rv), dp(p_ts), dp(p_tt), cd(resume.pop(0)))
            continue
        if done.value or it.value >= maxit:
            break
        v[:] = op(p)
        lk.kh_dot2(sz(n), dp(rhat), dp(v), None, dp(p_rv), None)
        if flex is None:
            lk.kh_bi_s(sz(n), dp(r), dp(v), dp(sv), dp(p_rv))
        else:
            lk.kh_bi_s_flex(sz(n), dp(r), dp(v), dp(sv), dp(p_rv), sz(flex[2]), dp(flex[0]), dp(flex[1]))
        t[:] = op(sv)
        lk.kh_dot2(sz(n), dp(t), dp(sv), dp(t), dp(p_ts), dp(p_tt))
        if flex is None:
            lk.kh_bi_xr(sz(n), dp(p), dp(sv), dp(t), dp(rhat), dp(y), dp(r), dp(p_rv), dp(p_ts), dp(p_tt), dp(p_rho),
                        dp(p_rr))
        else:
            lk.kh_bi_xr_flex(sz(n), dp(sv), dp(t), dp(rhat), dp(r), dp(p_ts), dp(p_tt), dp(p_rho), dp(p_rr), sz(flex[2]),
                             dp(flex[0]), dp(flex[1]))
        lk.kh_bi_p(sz(n), dp(r), dp(v), dp(p), dp(p_rv), dp(p_ts), dp(p_tt), dp(p_rho), dp(p_rr))
    return y, it.value, done.value


@pytest.mark.parametrize("k,nx", [(1, 6), (2, 4)])
def test_bicgstab_kernels_on_the_host(libs, k, nx):
    """the tentative-velocity solve of tests/test_tent_host.py with the engine's BiCGStab kernels in place of numpy"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    mesh, o, Q0, Qs, adt = _problem(k, nx, "upwind")
    ht = HostTentative(libs["tent"], mesh, k)
    lk = libs["krylov"]
    b = Q0 + 0.01 * np.random.default_rng(11).standard_normal(Q0.shape)
    M = sp.diags(np.repeat(o.detJ, o.nQ))
    x_ref = spla.spsolve((M - adt * o.f_impl_matrix(Qs)).tocsc(), M @ b.ravel()).reshape(b.shape)
    nq, nmu = 2 * ht.nq1 * ht.nc, ht.nm * ht.nf
    inv_aalpha = 1.0 / (adt * ht.alpha)
    Qstar, bs = soa(Qs), soa(b)

    def split(vec):
        return (np.ascontiguousarray(vec[:nq].reshape(2 * ht.nq1, ht.nc)),
                np.ascontiguousarray(vec[nq:].reshape(ht.nm, ht.nf)))

    tent_xh = np.zeros((2 * ht.nq1, ht.nc))  # what the operator application leaves behind: [Phat^-1 in]_x

    def op(vec):  # A_aug Phat^-1 (run_tentative_aug)
        vx, vmu = split(vec)
        mu, nyx = ht.precond_x(inv_aalpha, vx, vmu)
        xh = ht.xhat(vx, mu)
        tent_xh[:] = xh
        out_x = ht.fimpl(True, Qstar, xh, 1.0, -adt, Z=vx, alpha=0.0)
        out_mu = ht.sweep(inv_aalpha, nyx, mu, 0.0, 0.0, 0, 1)
        return np.concatenate([out_x.ravel(), out_mu.ravel()])

    r0 = np.concatenate([bs.ravel(), np.zeros(nmu)])
    y, its, done = bicgstab_kernels(lk, op, nq + nmu, r0, float(bs.ravel() @ bs.ravel()), 1e-12, 400)
    assert done == 1
    yx, ymu = split(y)
    mu, _ = ht.precond_x(inv_aalpha, yx, ymu)
    x = ht.xhat(yx, mu)
    assert np.abs(aos(x, o.nQ1) - x_ref).max() < 1e-9 * np.abs(x_ref).max()
    _, its_np = ht.solve(Qstar, adt, True, bs, 1e-12, False)
    assert abs(its - its_np) <= 2, (its, its_np)
    # flexible variant ("tent_flex"): x accumulated from the preconditioned directions, no recovery step
    xf = np.zeros((2 * ht.nq1, ht.nc))
    _, its_f, done = bicgstab_kernels(lk, op, nq + nmu, r0, float(bs.ravel() @ bs.ravel()), 1e-12, 400,
                                      flex=(tent_xh, xf, nq))
    assert done == 1 and abs(its_f - its) <= 2
    assert np.abs(aos(xf, o.nQ1) - x_ref).max() < 1e-9 * np.abs(x_ref).max()


def test_bicgstab_resume_continues_the_same_iteration(libs):
    """k_bi_resume (the acceptance loop of run_tentative_aug: a run that met its tolerance is continued with a tighter
    one): a solve stopped at 1e-5 and 1e-8 and resumed each time produces bit for bit the solution and the iteration
    count of the uninterrupted solve to 1e-12 -- the direction update skipped at the converged iteration is rebuilt
    from the partial sums that are still in place"""
    k, nx = 2, 4
    mesh, o, Q0, Qs, adt = _problem(k, nx, "upwind")
    ht = HostTentative(libs["tent"], mesh, k)
    lk = libs["krylov"]
    b = Q0 + 0.01 * np.random.default_rng(5).standard_normal(Q0.shape)
    nq, nmu = 2 * ht.nq1 * ht.nc, ht.nm * ht.nf
    inv_aalpha = 1.0 / (adt * ht.alpha)
    Qstar, bs = soa(Qs), soa(b)
    tent_xh = np.zeros((2 * ht.nq1, ht.nc))

    def op(vec):
        vx = np.ascontiguousarray(vec[:nq].reshape(2 * ht.nq1, ht.nc))
        vmu = np.ascontiguousarray(vec[nq:].reshape(ht.nm, ht.nf))
        mu, nyx = ht.precond_x(inv_aalpha, vx, vmu)
        xh = ht.xhat(vx, mu)
        tent_xh[:] = xh
        out_x = ht.fimpl(True, Qstar, xh, 1.0, -adt, Z=vx, alpha=0.0)
        out_mu = ht.sweep(inv_aalpha, nyx, mu, 0.0, 0.0, 0, 1)
        return np.concatenate([out_x.ravel(), out_mu.ravel()])

    r0 = np.concatenate([bs.ravel(), np.zeros(nmu)])
    bb = float(bs.ravel() @ bs.ravel())
    x_one = np.zeros((2 * ht.nq1, ht.nc))
    _, its_one, done = bicgstab_kernels(lk, op, nq + nmu, r0, bb, 1e-12, 400, flex=(tent_xh, x_one, nq))
    assert done == 1
    x_res = np.zeros((2 * ht.nq1, ht.nc))
    _, its_res, done = bicgstab_kernels(lk, op, nq + nmu, r0, bb, 1e-5, 400, flex=(tent_xh, x_res, nq),
                                        resume=(1e-3, 1e-4))
    assert done == 1 and its_res == its_one
    assert np.array_equal(x_res, x_one)


def test_fp32_stored_sweeps_need_and_work_with_the_flexible_update(libs):
    """``tent_fp32`` (k_tent_sweep32: iterate and correction of the Chebyshev sweeps stored in FP32): with the present
    recovery x = [Phat^-1 y]_x the attainable accuracy drops to ~1e-8, with the flexible update it is back at
    round-off -- which is why the engine only uses the FP32 sweeps together with ``tent_flex``"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    k, nx = 2, 4
    mesh, o, Q0, Qs, adt = _problem(k, nx, "upwind")
    lk, lt = libs["krylov"], libs["tent"]
    FP = ctypes.POINTER(ctypes.c_float)

    class HostTentative32(HostTentative):
        def schur_solve(self, inv_aalpha, t):  # tent_schur_solve, FP32 branch
            a, b_ = self.lmax / 8.0, 1.1 * self.lmax
            theta, delta = 0.5 * (b_ + a), 0.5 * (b_ - a)
            sigma = theta / delta
            rho, coefs = 1.0 / sigma, [(0.0, 1.0 / theta)]
            for _ in range(1, self.sweeps):
                rho_new = 1.0 / (2.0 * sigma - rho)
                coefs.append((rho_new * rho, 2.0 * rho_new / delta))
                rho = rho_new
            x, x2, d = (np.zeros((self.nm, self.nf), np.float32) for _ in range(3))
            out64 = np.zeros((self.nm, self.nf))
            for j, (cdv, crv) in enumerate(coefs):
                last = j == len(coefs) - 1
                assert lt.th_sweep32(self.k, self.nf, ip(self.facet_local), dp(self.tc), ip(self.tcol), ip(self.tbits),
                                     cd(inv_aalpha), dp(t), x.ctypes.data_as(FP), d.ctypes.data_as(FP),
                                     None if last else x2.ctypes.data_as(FP), dp(out64) if last else None, cd(cdv),
                                     cd(crv), int(j == 0)) == 0
                x, x2 = x2, x
            return out64

    b = Q0 + 0.01 * np.random.default_rng(11).standard_normal(Q0.shape)
    M = sp.diags(np.repeat(o.detJ, o.nQ))
    x_ref = spla.spsolve((M - adt * o.f_impl_matrix(Qs)).tocsc(), M @ b.ravel()).reshape(b.shape)
    ht = HostTentative32(lt, mesh, k)
    nq, nmu = 2 * ht.nq1 * ht.nc, ht.nm * ht.nf
    inv_aalpha = 1.0 / (adt * ht.alpha)
    Qstar, bs = soa(Qs), soa(b)
    tent_xh = np.zeros((2 * ht.nq1, ht.nc))

    def split(vec):
        return (np.ascontiguousarray(vec[:nq].reshape(2 * ht.nq1, ht.nc)),
                np.ascontiguousarray(vec[nq:].reshape(ht.nm, ht.nf)))

    def op(vec):
        vx, vmu = split(vec)
        mu, nyx = ht.precond_x(inv_aalpha, vx, vmu)
        tent_xh[:] = ht.xhat(vx, mu)
        out_x = ht.fimpl(True, Qstar, tent_xh, 1.0, -adt, Z=vx, alpha=0.0)
        out_mu = ht.sweep(inv_aalpha, nyx, mu, 0.0, 0.0, 0, 1)
        return np.concatenate([out_x.ravel(), out_mu.ravel()])

    r0, bb = np.concatenate([bs.ravel(), np.zeros(nmu)]), float(bs.ravel() @ bs.ravel())
    y, its, done = bicgstab_kernels(lk, op, nq + nmu, r0, bb, 1e-12, 400)
    yx, ymu = split(y)
    mu, _ = ht.precond_x(inv_aalpha, yx, ymu)
    err_recovered = np.abs(aos(ht.xhat(yx, mu), o.nQ1) - x_ref).max() / np.abs(x_ref).max()
    xf = np.zeros((2 * ht.nq1, ht.nc))
    _, its_f, done_f = bicgstab_kernels(lk, op, nq + nmu, r0, bb, 1e-12, 400, flex=(tent_xh, xf, nq))
    err_flex = np.abs(aos(xf, o.nQ1) - x_ref).max() / np.abs(x_ref).max()
    print(f"FP32-stored sweeps: iterations {its} / {its_f}, error against the direct solve: recovered {err_recovered:.1e}, "
          f"flexible {err_flex:.1e}")
    assert done == 1 and done_f == 1
    assert err_flex < 1e-10 < err_recovered
