"""Golden vectors (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py from the oracle).

CPU part: the oracle still reproduces them (regression pin; PARITY UNPINNED against the reference
itself).  GPU part: the CUDA engine, called through the C-ABI, reproduces them to
the 1e-10 relative tolerance BASELINE.json's north_star states.
"""
import os

import numpy as np
import pytest

from conftest import require_degree
from incompressibleeulerhdg_b200.mesh import UnitSquareMesh

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))
RTOL = 1e-10


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


# ---- CPU: oracle vs golden ------------------------------------------------------------------------
@pytest.mark.parametrize("k", [1, 2])
def test_oracle_poisson_golden(k):
    from oracle.hdg_oracle import HDGOracle

    m = UnitSquareMesh(4, perturb=0.15)
    g = lambda n: GOLD[f"poisson_k{k}/{n}"]
    Q, p, l = HDGOracle(m, k).solve_condensed(g("Ru"), g("Rp"), g("Rl"))
    assert rel(Q, g("Q")) < 1e-12 and rel(p, g("p")) < 1e-12 and rel(l, g("l")) < 1e-12


def test_oracle_chorin_golden():
    from oracle.timesteppers import ChorinOracle, TaylorGreenOracle

    m = UnitSquareMesh(4, perturb=0.1)
    Q, p = ChorinOracle(m, 2, 0.02).solve(TaylorGreenOracle("exponential", 0.5), 0.04)
    assert rel(Q, GOLD["chorin_k2/Q"]) < 1e-12 and rel(p, GOLD["chorin_k2/p"]) < 1e-12


def test_golden_config0_is_accurate():
    """BASELINE.json configs[0]: the stored fully implicit k=1 16x16 solution stays close to the exact
    stationary Taylor-Green velocity (`model_problems.py:56-66`) after 10 steps"""
    assert float(GOLD["implicit_k1_nx16/err_Q"]) < 5e-3


# ---- GPU: engine vs golden ------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 2])
def test_engine_poisson_golden(k):
    from incompressibleeulerhdg_b200.engine import HDGEngine

    require_degree(k)
    m = UnitSquareMesh(4, perturb=0.15)
    g = lambda n: GOLD[f"poisson_k{k}/{n}"]
    eng = HDGEngine(m, k)
    eng.setup_poisson()
    Q, p, l, its = eng.poisson_apply_host(g("Ru"), g("Rp"), g("Rl"), rtol=1e-13)
    assert its > 0
    assert rel(Q, g("Q")) < RTOL and rel(p, g("p")) < RTOL and rel(l, g("l")) < RTOL


@pytest.mark.gpu
def test_engine_chorin_golden():
    from incompressibleeulerhdg_b200 import timesteppers as TS
    from incompressibleeulerhdg_b200.model_problems import TaylorGreen

    require_degree(2)
    m = UnitSquareMesh(4, perturb=0.1)
    ts = TS.IncompressibleEulerHDGImplicit(m, 2, 0.02, krylov_rtol=1e-13)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q0, p0 = prob.initial_condition()
    Q, p = ts.solve(Q0, p0, None, prob.f_rhs(), 0.04)
    assert rel(Q.to_host(), GOLD["chorin_k2/Q"]) < RTOL and rel(p.to_host(), GOLD["chorin_k2/p"]) < RTOL


@pytest.mark.gpu
def test_engine_imex_golden():
    from incompressibleeulerhdg_b200 import timesteppers as TS
    from incompressibleeulerhdg_b200.model_problems import TaylorGreen

    require_degree(1)
    m = UnitSquareMesh(5, perturb=0.1)
    ts = TS.IncompressibleEulerHDGIMEXSSP2_332(m, 1, 0.02, n_richardson=2, krylov_rtol=1e-13)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q0, p0 = prob.initial_condition()
    Q, p = ts.solve(Q0, p0, None, prob.f_rhs(), 0.02)
    assert rel(Q.to_host(), GOLD["imex_ssp2_k1/Q"]) < RTOL and rel(p.to_host(), GOLD["imex_ssp2_k1/p"]) < RTOL


@pytest.mark.gpu
def test_engine_config0_fully_implicit_golden():
    """BASELINE.json configs[0]: HDG fully implicit, k=1, 16x16, stationary solution, 10 steps"""
    from incompressibleeulerhdg_b200 import timesteppers as TS
    from incompressibleeulerhdg_b200.model_problems import TaylorGreen

    require_degree(1)
    m = UnitSquareMesh(16)
    ts = TS.IncompressibleEulerHDGImplicit(m, 1, 0.1, use_projection_method=False, krylov_rtol=1e-13)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.0)  # kappa = 0: zero forcing
    Q0, p0 = prob.initial_condition()
    Q, p = ts.solve(Q0, p0, None, prob.f_rhs(), 1.0)
    assert rel(Q.to_host(), GOLD["implicit_k1_nx16/Q"]) < RTOL
    assert rel(p.to_host(), GOLD["implicit_k1_nx16/p"]) < 1e-9


# ---- passive tracer (tests/golden/golden_tracer_v1.npz, made by tests/golden/make_golden_tracer.py) ------
GOLD_T = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_tracer_v1.npz"))


def _tracer0(x, y):
    return np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y)


def test_oracle_tracer_golden():
    from oracle.hdg_oracle import HDGOracle
    from oracle.timesteppers import ChorinOracle, TaylorGreenOracle
    from oracle.tracer import TracerOracle

    t = TracerOracle(HDGOracle(UnitSquareMesh(4, perturb=0.15), 2))
    U = t.project_cg(GOLD_T["project_k2/Q"])
    assert rel(U, GOLD_T["project_k2/U"]) < 1e-12
    assert rel(t.advection(GOLD_T["advect_k2/q"], U), GOLD_T["advect_k2/adv"]) < 1e-12
    orc = ChorinOracle(UnitSquareMesh(4, perturb=0.1), 2, 0.02)
    orc.solve(TaylorGreenOracle("exponential", 0.5), 0.04, q_initial=_tracer0)
    assert rel(orc.q_tracer, GOLD_T["chorin_k2/q"]) < 1e-12


@pytest.mark.gpu
def test_engine_tracer_golden():
    from incompressibleeulerhdg_b200.engine import HDGEngine
    from incompressibleeulerhdg_b200.functions import Expression
    from incompressibleeulerhdg_b200.model_problems import TaylorGreen
    from incompressibleeulerhdg_b200.timesteppers import (IncompressibleEulerHDGIMEXSSP2_332,
                                                          IncompressibleEulerHDGImplicit)

    require_degree(2)
    eng = HDGEngine(UnitSquareMesh(4, perturb=0.15), 2)
    eng.tracer_setup()
    dU, dadv = eng.empty(0), eng.empty(1)
    eng.project_cg_dev(eng.upload(0, GOLD_T["project_k2/Q"]), dU, rtol=1e-14)
    assert rel(eng.download(0, dU), GOLD_T["project_k2/U"]) < RTOL
    eng.tracer_advection_dev(dU, eng.upload(1, GOLD_T["advect_k2/q"]), dadv)
    assert rel(eng.download(1, dadv), GOLD_T["advect_k2/adv"]) < RTOL
    ts = IncompressibleEulerHDGImplicit(UnitSquareMesh(4, perturb=0.1), 2, 0.02, krylov_rtol=1e-13)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    ts.solve(*prob.initial_condition(), Expression(_tracer0, 0), prob.f_rhs(), 0.04)
    assert rel(ts.q_tracer.to_host(), GOLD_T["chorin_k2/q"]) < RTOL
    require_degree(1)
    ts = IncompressibleEulerHDGIMEXSSP2_332(UnitSquareMesh(5, perturb=0.1), 1, 0.02, n_richardson=2, krylov_rtol=1e-13)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    ts.solve(*prob.initial_condition(), Expression(_tracer0, 0), prob.f_rhs(), 0.02)
    assert rel(ts.q_tracer.to_host(), GOLD_T["imex_ssp2_k1/q"]) < RTOL


# ---- IMEX tableaux: golden data produced by EXECUTING the reference's own property bodies ------------------------
# (tests/golden/make_golden_tableaux.py; src/timesteppers/hdg_imex.py:668-1038).  This is the one part of the path
# that is pinned against the reference itself rather than against the restatement.
def _tableaux_golden():
    import json

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tableaux_v1.json")
    return json.load(open(path))["classes"]


ORACLE_KEY = {"IncompressibleEulerHDGIMEXImplicit": "imex_implicit", "IncompressibleEulerHDGIMEXARS2_232": "imex_ars2_232",
              "IncompressibleEulerHDGIMEXARS3_443": "imex_ars3_443", "IncompressibleEulerHDGIMEXSSP2_332": "imex_ssp2_332",
              "IncompressibleEulerHDGIMEXSSP3_433": "imex_ssp3_433"}


@pytest.mark.parametrize("cls_name", sorted(ORACLE_KEY))
def test_imex_tableaux_equal_the_reference(cls_name):
    import incompressibleeulerhdg_b200.timesteppers as TS
    from oracle.timesteppers import TABLEAUX

    gold = _tableaux_golden()[cls_name]
    cls = getattr(TS, cls_name)
    orc = TABLEAUX[ORACLE_KEY[cls_name]]
    assert cls.nstages.fget(None) == gold["nstages"] == orc["nstages"]
    for prop in ("_a_expl", "_a_impl", "_b_expl", "_b_impl", "_c_expl"):
        g = np.asarray(gold[prop], dtype=float)
        mine = np.asarray(getattr(cls, prop).fget(None), dtype=float)
        assert mine.shape == g.shape and np.array_equal(mine, g), (cls_name, prop)  # bit-exact
        assert np.array_equal(np.asarray(orc[prop[1:]], dtype=float), g), (cls_name, prop, "oracle")


def test_imex_tableaux_labels_and_order_conditions():
    """what the reference's tableau data does and does not satisfy -- the quirks are part of the
    behaviour to reproduce, so they are pinned here rather than "fixed":
      * ARS3(4,4,3) stores SIX implicit weights for five stages; the loops read the first five,
        [0, 3/2, -3, 2, 1/2], which sum to one but give b_impl . c = 1/4 instead of 1/2;
      * SSP2(3,3,2) evaluates the forcing at c_expl = (0, 1, 1/2) although the row sums of A_expl are
        (0, 1/2, 1)."""
    import incompressibleeulerhdg_b200.timesteppers as TS

    gold = _tableaux_golden()
    assert len(gold) == 5
    for name, g in gold.items():
        s_ = g["nstages"]
        ae, ai = np.asarray(g["_a_expl"]), np.asarray(g["_a_impl"])
        be, bi, ce = np.asarray(g["_b_expl"]), np.asarray(g["_b_impl"]), np.asarray(g["_c_expl"])
        assert ae.shape == ai.shape == (s_, s_) and be.shape == ce.shape == (s_,)
        assert bi.shape == ((6,) if name.endswith("ARS3_443") else (s_,))
        bi = bi[:s_]
        assert abs(be.sum() - 1) < 1e-14 and abs(bi.sum() - 1) < 1e-14  # consistency (first order)
        assert np.allclose(np.triu(ae), 0) and np.allclose(np.triu(ai, 1), 0)  # explicit / diagonally implicit
        assert np.allclose(ae.sum(axis=1), ce, atol=1e-14) == (not name.endswith("SSP2_332"))
        if "Implicit" in name:
            continue
        ci = ai.sum(axis=1)
        second = [abs(b_ @ c_ - 0.5) < 1e-13 for b_ in (be, bi) for c_ in (ce, ci)]
        assert second == ([True, True, False, False] if name.endswith("ARS3_443") else [True] * 4), (name, second)
    # the label is part of the driver's output (`driver.py:305`); instantiating needs a GPU, so compare the
    # string baked into the factory call instead
    import inspect

    src = inspect.getsource(TS.hdg_imex)
    for g in gold.values():
        assert f'"{g["label"]}"' in src, g["label"]


# ---- host-side pieces pinned against the reference itself (tests/golden/make_golden_reference_host.py) ----------
def _host_golden():
    import json

    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_host_v1.json")))


def test_log_summary_text_equals_the_reference():
    import io

    from incompressibleeulerhdg_b200.auxilliary.logging import PerformanceLog, log_summary

    gold = _host_golden()
    saved = {k: list(v) for k, v in PerformanceLog.records.items()}
    try:
        PerformanceLog.reset()
        out = io.StringIO()
        assert log_summary(file=out) == [] and out.getvalue() == ""  # the reference prints nothing without timers
        for label, ts in gold["timings"].items():
            PerformanceLog.records[label].extend(ts)
        log_summary(file=out)
        assert out.getvalue() == gold["log_summary"]
    finally:
        PerformanceLog.reset()
        for k, v in saved.items():
            PerformanceLog.records[k].extend(v)


def test_averager_equals_the_reference():
    from incompressibleeulerhdg_b200.auxilliary.utils import Averager

    gold = _host_golden()
    a = Averager()
    for x, (n, v) in zip(gold["samples"], gold["averager"]["running"]):
        a.update(x)
        assert a.n_samples == n and a.value == v  # same recurrence => bit-identical
    assert repr(a) == gold["averager"]["repr"]
    a.reset()
    assert [a.n_samples, float(a.value)] == gold["averager"]["after_reset"]


def test_shear_flow_pressure_uses_the_reference_fourier_coefficients():
    """`model_problems.py:166-187`: p_0 = delta cos(x) sum_k c_k sin((2k+1)(y - pi)) / (1 + (2k+1)^2) with the
    reference's own quad() coefficients c_k"""
    from incompressibleeulerhdg_b200.model_problems import DoubleLayerShearFlow

    g = _host_golden()["shear"]
    prob = DoubleLayerShearFlow(None, None)
    assert prob.rho == g["rho"] and prob.delta == g["delta"]
    rng = np.random.default_rng(3)
    x, y = rng.uniform(0, 2 * np.pi, 50), rng.uniform(0, 2 * np.pi, 50)
    n = 2 * np.arange(g["kmax"]) + 1
    series = (np.asarray(g["fourier_coefficient"])[:, None] * np.sin(n[:, None] * (y[None, :] - np.pi))
              / (1 + n[:, None] ** 2)).sum(axis=0)
    _, p0 = prob.initial_condition()
    assert np.abs(p0(x, y) - g["delta"] * np.cos(x) * series).max() < 1e-13


# ---- mid-size golden (tests/golden/golden_midsize_v1.npz, made once by tests/golden/make_golden_midsize.py) ----------
GOLD_M = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_midsize_v1.npz"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfl032", "cfl32"])
def test_engine_chorin_midsize_golden(name):
    """BASELINE.json configs[2] reduced to nx = 32 (2048 cells), k = 2, two Chorin steps, at the bench's CFL and at ten
    times it: the engine against the oracle's committed sparse-direct result, 1e-10"""
    from incompressibleeulerhdg_b200 import timesteppers as TS
    from incompressibleeulerhdg_b200.model_problems import TaylorGreen

    require_degree(2)
    dt = float(GOLD_M[f"chorin_k2_nx32_{name}/dt"])
    m = UnitSquareMesh(32, perturb=0.1)
    ts = TS.IncompressibleEulerHDGImplicit(m, 2, dt, krylov_rtol=1e-13)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q, p = ts.solve(*prob.initial_condition(), None, prob.f_rhs(), 2 * dt)
    assert rel(Q.to_host(), GOLD_M[f"chorin_k2_nx32_{name}/Q"]) < 1e-10
    assert rel(p.to_host(), GOLD_M[f"chorin_k2_nx32_{name}/p"]) < 1e-10


def test_cpu_baseline_midsize_golden():
    """the compiled CPU baseline of bench.py (oracle/cpu_ref) against the same committed vectors"""
    from oracle.cpu_ref import ChorinCpuRef
    from oracle.timesteppers import TaylorGreenOracle

    dt = float(GOLD_M["chorin_k2_nx32_cfl032/dt"])
    c = ChorinCpuRef(UnitSquareMesh(32, perturb=0.1), 2, dt, rtol=1e-13)
    Q, p = c.solve(TaylorGreenOracle("exponential", 0.5), 2 * dt)
    assert rel(Q, GOLD_M["chorin_k2_nx32_cfl032/Q"]) < 1e-10 and rel(p, GOLD_M["chorin_k2_nx32_cfl032/p"]) < 1e-10
