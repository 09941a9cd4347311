"""Knobs of the tentative-velocity solve, all on by default since the round-2 A/B on a B200
(profiles/r2/bench_r2a_knob_ab.jsonl): ``tent_cellblock`` (csrc/hdg_advblock.cuh, the cell-block advection
preconditioner), ``tent_scaledx`` (scaled facet Schur complement, csrc/hdg_tent.cuh), ``tent_flex`` (flexible solution
update of BiCGStab, csrc/hdg_krylov.cuh) and ``tent_fp32`` (FP32-stored Schur sweep vectors).  Every combination must
leave the solution where the oracle has it."""
import os

import numpy as np
import pytest

import incompressibleeulerhdg_b200.timesteppers as TS
from conftest import require_degree
from incompressibleeulerhdg_b200.mesh import UnitSquareMesh
from incompressibleeulerhdg_b200.model_problems import TaylorGreen
from oracle.timesteppers import ChorinOracle, TaylorGreenOracle

pytestmark = [pytest.mark.gpu]


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


# (k = 3 with the central flux is left out: at this CFL number the solver *without* the cell blocks needs more than
# 600 iterations in the host port of tests/test_tent_host.py -- 122 with them -- so there is no baseline to compare)
@pytest.mark.parametrize("k,nx,flux", [(1, 8, "upwind"), (1, 8, "centered"), (2, 8, "upwind"), (2, 8, "centered"),
                                       (3, 4, "upwind")])
def test_cellblock_preconditioner_keeps_the_solution_and_cuts_iterations(k, nx, flux):
    require_degree(k)
    mesh, dt, nt = UnitSquareMesh(nx, perturb=0.1), 0.32 / nx, 3
    runs = {}
    for on in (0, 1):
        ts = TS.IncompressibleEulerHDGImplicit(mesh, k, dt, flux=flux, krylov_rtol=1e-13, warm_start=False)
        ts.engine.set_tuning("tent_cellblock", on)
        prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
        Q, p = ts.solve(*prob.initial_condition(), None, prob.f_rhs(), nt * dt)
        runs[on] = (Q.to_host(), p.to_host(), ts.niter_tentative.value)
    Qo, po = ChorinOracle(mesh, k, dt, flux=flux).solve(TaylorGreenOracle("exponential", 0.5), nt * dt)
    print(f"k={k} {flux}: BiCGStab iterations per solve {runs[0][2]:.1f} -> {runs[1][2]:.1f}; "
          f"knob on vs oracle: velocity {rel(runs[1][0], Qo):.2e} pressure {rel(runs[1][1], po):.2e}")
    assert rel(runs[1][0], Qo) < 1e-10 and rel(runs[1][1], po) < 1e-10
    assert rel(runs[1][0], runs[0][0]) < 1e-10
    assert runs[1][2] < 0.8 * runs[0][2]


@pytest.mark.parametrize("knobs", [("tent_flex",), ("tent_flex", "tent_cellblock"), ("tent_flex", "tent_fp32"),
                                   ("tent_flex", "tent_fp32", "tent_cellblock"), ("tent_scaledx",),
                                   ("tent_scaledx", "tent_flex", "tent_fp32")])
def test_flexible_bicgstab_update_keeps_the_solution(knobs):
    """``tent_flex``: the tentative velocity is accumulated from the preconditioned directions (k_bi_s_flex /
    k_bi_xr_flex, checked on the CPU in tests/test_krylov_host.py) instead of being recovered from the accumulated
    Krylov vector; same solution, about the same iteration count, with and without a warm start.  ``tent_fp32`` on top
    of it stores the vectors of the Chebyshev sweeps in FP32 (k_tent_sweep32)"""
    k, nx = 2, 8
    require_degree(k)
    mesh, dt, nt = UnitSquareMesh(nx, perturb=0.1), 0.32 / nx, 3
    Qo, po = ChorinOracle(mesh, k, dt).solve(TaylorGreenOracle("exponential", 0.5), nt * dt)
    for warm in (False, True):
        its = {}
        for on in (0, 1):
            ts = TS.IncompressibleEulerHDGImplicit(mesh, k, dt, krylov_rtol=1e-13, warm_start=warm)
            for name in knobs:
                ts.engine.set_tuning(name, on)
            prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
            Q, p = ts.solve(*prob.initial_condition(), None, prob.f_rhs(), nt * dt)
            its[on] = ts.niter_tentative.value
            assert rel(Q.to_host(), Qo) < 1e-10 and rel(p.to_host(), po) < 1e-10, (knobs, warm, on)
        print(f"{knobs} warm_start={warm}: BiCGStab iterations per solve {its[0]:.1f} -> {its[1]:.1f}")
        assert its[1] <= 1.15 * its[0] + 3


@pytest.mark.parametrize("k,nx,flux", [(1, 8, "upwind"), (2, 8, "upwind"), (2, 9, "upwind"), (2, 9, "centered"),
                                       (3, 4, "upwind")])
def test_operator_variants_of_the_iteration_give_the_same_solution(k, nx, flux):
    """``fimpl_split``: the penalty-free operator of the augmented Krylov iteration as k_fimpl (0), k_fimpl_c (1, default:
    one thread per (cell, component), csrc/hdg_flow.cuh) or k_fimpl_t (3: the same with the cell's rows staged in shared
    memory by TMA bulk copies; k <= 2, full 64-cell tiles, the last partial tile and k = 3 take the global path) and
    ``sweep_minblocks`` (register-allocation variants of the Schur sweeps): the same operator up to the order of the
    floating-point operations, so the same solution and about the same iteration counts (the central flux needs ~90
    iterations here, where round-off moves the count by a few); nx = 9 gives 162 cells = 2 staged tiles + a partial one"""
    require_degree(k)
    mesh, dt, nt = UnitSquareMesh(nx, perturb=0.1), 0.32 / nx, 2
    Qo, po = ChorinOracle(mesh, k, dt, flux=flux).solve(TaylorGreenOracle("exponential", 0.5), nt * dt)
    its = {}
    for split, minb in ((0, 5), (1, 6), (3, 8)):
        ts = TS.IncompressibleEulerHDGImplicit(mesh, k, dt, flux=flux, krylov_rtol=1e-13, warm_start=False)
        ts.engine.set_tuning("fimpl_split", split)
        ts.engine.set_tuning("sweep_minblocks", minb)
        prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
        Q, p = ts.solve(*prob.initial_condition(), None, prob.f_rhs(), nt * dt)
        its[split] = ts.niter_tentative.value
        counts = ts.engine.kernel_counts()
        assert rel(Q.to_host(), Qo) < 1e-10 and rel(p.to_host(), po) < 1e-10, (split, minb)
        expected = {0: "k_fimpl", 1: "k_fimpl_c", 3: "k_fimpl_t" if k <= 2 else "k_fimpl_c"}[split]
        assert counts.get(expected, 0) > 0, (split, sorted(counts))
    print(f"k={k} nx={nx} {flux}: BiCGStab iterations per solve {its}")
    assert max(its.values()) - min(its.values()) <= 0.1 * max(its.values()) + 2.0, its
