"""GPU parity of the passive-tracer path (`common.py:110-129`, `hdg_implicit.py:73-96,
192-193`, `hdg_imex.py:415-448,622-623,638-639`) against the oracle, through the C-ABI."""
import numpy as np
import pytest

from incompressibleeulerhdg_b200.engine import HDGEngine
from incompressibleeulerhdg_b200.functions import Expression
from incompressibleeulerhdg_b200.mesh import PeriodicSquareMesh, UnitDiskMesh, UnitSquareMesh
from incompressibleeulerhdg_b200.model_problems import TaylorGreen
from incompressibleeulerhdg_b200.timesteppers import (IncompressibleEulerHDGIMEXARS2_232,
                                                      IncompressibleEulerHDGIMEXSSP2_332,
                                                      IncompressibleEulerHDGImplicit)
from oracle.hdg_oracle import HDGOracle
from oracle.timesteppers import ChorinOracle, IMEXOracle, TaylorGreenOracle
from oracle.tracer import TracerOracle

pytestmark = pytest.mark.gpu

TOL = 1e-10  # relative, FP64 (north_star)


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def tracer0(x, y):
    return np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y)


def tg_velocity(x, y):
    return (-np.cos((x - 0.5) * np.pi) * np.sin((y - 0.5) * np.pi), np.sin((x - 0.5) * np.pi) * np.cos((y - 0.5) * np.pi))


@pytest.fixture(params=[1, 2, 3, 4])
def k(request):
    from incompressibleeulerhdg_b200.engine import load_library

    if not (load_library().hdg_supported_degrees() >> request.param) & 1:
        pytest.skip("degree not compiled in")
    return request.param


@pytest.mark.parametrize("mesh_fn", [lambda: UnitSquareMesh(5, perturb=0.15), lambda: UnitDiskMesh(1)])
def test_project_cg_matches_oracle(k, mesh_fn):
    m = mesh_fn()
    o, eng = HDGOracle(m, k), HDGEngine(m, k)
    t = TracerOracle(o)
    eng.tracer_setup()
    Q = o.interpolate_cell(tg_velocity, "Q") + 0.1 * np.random.default_rng(2).standard_normal((m.nc, 2, o.nQ1))
    dU = eng.empty(0)
    its = eng.project_cg_dev(eng.upload(0, Q), dU, rtol=1e-14)
    U, Uo = eng.download(0, dU), t.project_cg(Q)
    print(f"k={k} nc={m.nc} ncg={eng.cg_space.ndof} pcg its={its} rel.err={rel(U, Uo):.2e}")
    assert 0 < its < 200
    assert rel(U, Uo) < TOL
    # a projection: applying it twice changes nothing
    dU2 = eng.empty(0)
    eng.project_cg_dev(dU, dU2, rtol=1e-14)
    assert rel(eng.download(0, dU2), U) < TOL


@pytest.mark.parametrize("mesh_fn", [lambda: UnitSquareMesh(5, perturb=0.15), lambda: PeriodicSquareMesh(4, L=2 * np.pi),
                                     lambda: UnitDiskMesh(1)])
def test_tracer_advection_matches_oracle(k, mesh_fn):
    m = mesh_fn()
    o, eng = HDGOracle(m, k), HDGEngine(m, k)
    t = TracerOracle(o)
    eng.tracer_setup(nq_facet=o.nq_facet)
    rng = np.random.default_rng(4)
    U = o.interpolate_cell(tg_velocity, "Q") + 0.1 * rng.standard_normal((m.nc, 2, o.nQ1))
    q = rng.standard_normal((m.nc, o.np_))
    acc = rng.standard_normal((m.nc, o.np_))
    adv = t.advection(q, U)
    dU, dq, dout = eng.upload(0, U), eng.upload(1, q), eng.empty(1)
    # 1: compile-time tables (default facet rule), 0: runtime tables
    for variant in (1, 0):
        eng.set_tuning("tracer_tables", variant)
        dout.zero_()
        eng.tracer_advection_dev(dU, dq, dout)
        err = rel(eng.download(1, dout), adv)
        print(f"k={k} {m.name} variant={variant} rel.err={err:.2e}")
        assert err < TOL
        dacc = eng.upload(1, acc)
        eng.tracer_advection_dev(dU, dq, dacc, c0=0.5, acc=dacc, c1=-0.25)  # acc aliases out
        assert rel(eng.download(1, dacc), 0.5 * acc - 0.25 * adv) < TOL


def test_tracer_advection_other_facet_rule():
    """a facet rule that is not the compiled-in default goes through the runtime-table kernel"""
    k = 2
    m = UnitSquareMesh(5, perturb=0.15)
    o = HDGOracle(m, k, nq_facet=7)
    eng = HDGEngine(m, k)
    eng.tracer_setup(nq_facet=7)
    rng = np.random.default_rng(6)
    U = o.interpolate_cell(tg_velocity, "Q") + 0.1 * rng.standard_normal((m.nc, 2, o.nQ1))
    q = rng.standard_normal((m.nc, o.np_))
    dout = eng.empty(1)
    eng.tracer_advection_dev(eng.upload(0, U), eng.upload(1, q), dout)
    assert rel(eng.download(1, dout), TracerOracle(o).advection(q, U)) < TOL


def test_tracer_requires_setup():
    from incompressibleeulerhdg_b200.engine import HDGError

    m = UnitSquareMesh(3)
    eng = HDGEngine(m, 1)
    with pytest.raises(HDGError):
        eng.project_cg_dev(eng.zeros(0), eng.empty(0))


@pytest.mark.parametrize("k,nx", [(1, 6), (2, 5)])
def test_chorin_tracer_matches_oracle(k, nx):
    from conftest import require_degree

    require_degree(k)
    mesh, dt, nt = UnitSquareMesh(nx, perturb=0.1), 0.02, 3
    ts = IncompressibleEulerHDGImplicit(mesh, k, dt, krylov_rtol=1e-13)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q0, p0 = prob.initial_condition()
    Q, p = ts.solve(Q0, p0, Expression(tracer0, 0), prob.f_rhs(), nt * dt)
    orc = ChorinOracle(mesh, k, dt)
    Qo, po = orc.solve(TaylorGreenOracle("exponential", 0.5), nt * dt, q_initial=tracer0)
    eq = rel(ts.q_tracer.to_host(), orc.q_tracer)
    print(f"k={k} tracer rel.err={eq:.2e} velocity {rel(Q.to_host(), Qo):.2e} cg-projection its={ts.niter_cg_projection.value:.1f}")
    assert rel(Q.to_host(), Qo) < TOL and rel(p.to_host(), po) < TOL
    assert eq < TOL


@pytest.mark.parametrize("name,cls", [("imex_ssp2_332", IncompressibleEulerHDGIMEXSSP2_332),
                                      ("imex_ars2_232", IncompressibleEulerHDGIMEXARS2_232)])
def test_imex_tracer_matches_oracle(name, cls):
    from conftest import require_degree

    k = 1
    require_degree(k)
    mesh, dt, nt = UnitSquareMesh(5, perturb=0.1), 0.02, 2
    ts = cls(mesh, k, dt, use_projection_method=True, krylov_rtol=1e-13)
    prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
    Q0, p0 = prob.initial_condition()
    Q, p = ts.solve(Q0, p0, Expression(tracer0, 0), prob.f_rhs(), nt * dt)
    orc = IMEXOracle(mesh, k, dt, tableau=name)
    Qo, po = orc.solve(TaylorGreenOracle("exponential", 0.5), nt * dt, q_initial=tracer0)
    eq = rel(ts.q_tracer.to_host(), orc.q_tracer)
    print(f"{name} tracer rel.err={eq:.2e} velocity {rel(Q.to_host(), Qo):.2e}")
    assert rel(Q.to_host(), Qo) < TOL
    assert eq < TOL


def test_tracer_mass_budget_large():
    """size-independent property at a size the oracle cannot reach: over one explicit Euler step
    int q dx changes by exactly dt int q div(u_cg) dx (chi = 1 has no facet jump), which is tiny for the
    nearly solenoidal projected Taylor-Green velocity"""
    nx, k, dt = 128, 2, 1e-3
    mesh = UnitSquareMesh(nx)
    eng = HDGEngine(mesh, k)
    eng.tracer_setup()
    from incompressibleeulerhdg_b200.functions import FunctionSpace

    V_Q, V_q = FunctionSpace(eng, "Q"), FunctionSpace(eng, "p")
    Qf = V_Q.interpolate(Expression(tg_velocity, 1))
    q = V_q.interpolate(Expression(lambda x, y: 1.0 + tracer0(x, y), 0))
    one = V_q.interpolate(Expression(lambda x, y: 1.0 + 0 * x, 0))
    U = eng.empty(0)
    its = eng.project_cg_dev(Qf.data, U, rtol=1e-13)
    out = eng.empty(1)
    eng.tracer_advection_dev(U, q.data, out, c0=1.0, acc=q.data, c1=dt)
    m0 = eng.l2_inner_dev(1, q.data, one.data)
    m1 = eng.l2_inner_dev(1, out, one.data)
    print(f"nx={nx} ncg={eng.cg_space.ndof} pcg its={its} mass {m0:.15f} -> {m1:.15f}")
    assert 0 < its < 200
    assert abs(m0 - 1.0) < 1e-5
    assert abs(m1 - m0) < 1e-4 * dt
