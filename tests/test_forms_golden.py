"""The oracle's hand-written forms against the reference's OWN form functions.

tests/golden/forms_v1.npz holds dual vectors obtained by *executing* `_f_impl`, `_pressure_gradient`, `_Gamma`,
`_weak_divergence`, the `a_mixed_poisson` expression (`src/timesteppers/hdg_imex.py:123-127,313-365`) and
`_tracer_advection` (`src/timesteppers/common.py:110-129`), cut out of the reference source with `ast`, through the
mini-UFL interpreter `oracle/miniufl.py` (tests/golden/make_golden_forms.py).  Every sign, factor, restriction and
measure therefore comes from the reference text; the oracle (`oracle/hdg_oracle.py`, `oracle/tracer.py`) must
reproduce the vectors to round-off.  The GPU parity tests compare the engine with the same oracle functions, which
closes the chain  reference forms == oracle == engine."""
import os

import numpy as np
import pytest

from incompressibleeulerhdg_b200.mesh import PeriodicSquareMesh, UnitSquareMesh
from oracle.hdg_oracle import HDGOracle
from oracle.tracer import TracerOracle

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "forms_v1.npz"))
CASES = {"square_k1": (lambda: UnitSquareMesh(3, perturb=0.15), 1), "square_k2": (lambda: UnitSquareMesh(3, perturb=0.15), 2),
         "periodic_k2": (lambda: PeriodicSquareMesh(3, L=2 * np.pi), 2)}
TOL = 1e-13


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(params=sorted(CASES))
def case(request):
    mesh_fn, k = CASES[request.param]
    tag = request.param
    return mesh_fn(), k, (lambda name: GOLD[f"{tag}/{name}"])


@pytest.mark.parametrize("flux", ["upwind", "centered"])
def test_f_impl(case, flux):
    """`_f_impl` hdg_imex.py:313-331, with a BDM-projected and with a raw (normal-discontinuous) Q*"""
    mesh, k, g = case
    o = HDGOracle(mesh, k, flux=flux)
    Q, Qs_raw = g("in_Q"), g("in_Qstar_raw")
    assert rel(o.f_impl_apply(Q, o.project_bdm(Qs_raw)), g(f"f_impl_{flux}")) < TOL
    assert rel(o.f_impl_apply(Q, Qs_raw), g(f"f_impl_{flux}_rawQstar")) < TOL


def test_pressure_gradient_and_weak_divergence(case):
    """`_pressure_gradient` :333-340, `_weak_divergence` :353-365"""
    mesh, k, g = case
    o = HDGOracle(mesh, k)
    assert rel(o.pressure_gradient(g("in_p"), g("in_lam")), g("pressure_gradient")) < TOL
    assert rel(o.weak_divergence(g("in_Q")), g("weak_divergence")) < TOL


def test_mixed_poisson_operator_and_gamma(case):
    """the monolithic matrix the condensation starts from == action of `a_mixed_poisson` (:123-127); its (psi, mu)
    rows == `_Gamma` (:342-351)"""
    mesh, k, g = case
    o = HDGOracle(mesh, k)
    Q, p, lam = g("in_Q"), g("in_p"), g("in_lam")
    K, (offp, offl, N) = o.assemble_monolithic()
    y = K @ np.concatenate([Q.ravel(), p.ravel(), lam.ravel()])
    assert rel(y[:offp].reshape(Q.shape), g("mixed_poisson_Q")) < TOL
    assert rel(y[offp:offl].reshape(p.shape), g("mixed_poisson_P")) < TOL
    assert rel(y[offl:].reshape(lam.shape), g("mixed_poisson_T")) < TOL
    assert rel(y[offp:offl].reshape(p.shape), g("Gamma_P")) < TOL
    assert rel(y[offl:].reshape(lam.shape), g("Gamma_T")) < TOL
    # and the same through the local blocks the condensation uses (S_K = D - C A^-1 B is built from them)
    b = o.local_blocks()
    td = o.trace_dofs()
    u = Q.reshape(mesh.nc, -1)
    lamK = lam.ravel()[td]
    Rp = np.einsum("nai,ni->na", b["B"], u) + np.einsum("nab,nb->na", b["T"], p) - o.tau * np.einsum("nla,nl->na", b["F"], lamK)
    assert rel(Rp, g("Gamma_P")) < TOL
    RlK = np.einsum("nli,ni->nl", b["E"], u) + o.tau * np.einsum("nla,na->nl", b["F"], p) - o.tau * np.einsum("nlm,nm->nl", b["G"], lamK)
    Rl = np.zeros(mesh.nf * o.nl1)
    np.add.at(Rl, td.ravel(), RlK.ravel())
    assert rel(Rl.reshape(lam.shape), g("Gamma_T")) < TOL


def test_tracer_advection(case):
    """`_tracer_advection` common.py:110-129 on a continuous velocity (what the form sees after the CG projection)"""
    mesh, k, g = case
    o = HDGOracle(mesh, k)
    t = TracerOracle.__new__(TracerOracle)  # the advection form needs no CG space
    t.o = o
    assert rel(t.advection(g("in_q"), g("in_Ucont")) * o.detJ[:, None], g("tracer_advection")) < TOL


CHORIN = {"chorin_square_k2": (lambda: UnitSquareMesh(3, perturb=0.15), 2),
          "chorin_periodic_k1": (lambda: PeriodicSquareMesh(3, L=2 * np.pi), 1)}


@pytest.mark.parametrize("tag", sorted(CHORIN))
def test_chorin_inline_forms(tag):
    """the forms written inline in `IncompressibleEulerHDGImplicit.solve` (`hdg_implicit.py:103-145`) against the
    operators the Chorin oracle assembles: tentative matrix M - dt f_impl (both fluxes), its right-hand side, the
    mixed-Poisson operator written out a second time, and the Poisson right-hand side -(1/dt) psi div(Q~) dx"""
    from oracle.timesteppers import ChorinOracle

    mesh_fn, k = CHORIN[tag]
    mesh = mesh_fn()
    g = lambda name: GOLD[f"{tag}/{name}"]
    dt = float(g("in_dt"))
    X, Q, f, Qs, p, lam = (g("in_" + n) for n in ("X", "Q", "f", "Qstar", "p", "lam"))
    for flux in ("upwind", "centered"):
        orc = ChorinOracle(mesh, k, dt, flux=flux)
        A = orc.tentative_matrix(Qs, dt)
        assert rel((A @ X.ravel()).reshape(X.shape), g(f"a_tentative_{flux}")) < TOL
    o = orc.o
    assert rel(orc.mass(Q) + dt * orc.mass(f), g("b_rhs_tentative")) < TOL
    K, (offp, offl, N) = o.assemble_monolithic()
    y = K @ np.concatenate([X.ravel(), p.ravel(), lam.ravel()])
    assert rel(y[:offp].reshape(X.shape), g("a_poisson_Q")) < TOL
    assert rel(y[offp:offl].reshape(p.shape), g("a_poisson_P")) < TOL
    assert rel(y[offl:].reshape(lam.shape), g("a_poisson_T")) < TOL
    assert rel(-(1.0 / dt) * o.cell_divergence(X), g("b_rhs_poisson")) < TOL


IMEX_CLASSES = {"IncompressibleEulerHDGIMEXImplicit": "imex_implicit", "IncompressibleEulerHDGIMEXARS2_232": "imex_ars2_232",
                "IncompressibleEulerHDGIMEXARS3_443": "imex_ars3_443", "IncompressibleEulerHDGIMEXSSP2_332": "imex_ssp2_332",
                "IncompressibleEulerHDGIMEXSSP3_433": "imex_ssp3_433"}


@pytest.mark.parametrize("cls_name", sorted(IMEX_CLASSES))
def test_imex_stage_forms(cls_name):
    """`_residual` / `_final_residual` (`hdg_imex.py:367-413`, with the recursion and its quirks executed from the
    reference) and the stage forms of `__init__` (`:177-179,233-247`) against what `IMEXOracle.step` assembles"""
    from oracle.timesteppers import IMEXOracle

    mesh, k = UnitSquareMesh(3, perturb=0.15), 1
    g = lambda name: GOLD[f"imex_{cls_name}/{name}"]
    dt = float(g("in_dt"))
    orc = IMEXOracle(mesh, k, dt, tableau=IMEX_CLASSES[cls_name])
    o, s_ = orc.o, orc.nstages
    for j in range(s_):
        orc.stage[j] = dict(Q=g(f"in_stage{j}_Q"), p=g(f"in_stage{j}_p"), l=g(f"in_stage{j}_l"))
        orc.b_rhs[j] = g(f"in_b_rhs{j}")
    X = g("in_X")
    assert rel(orc.final_residual(), g("final_residual")) < TOL
    for i in range(1, s_):
        a = orc.a_impl[i, i]
        Qstar, st = g(f"in_Qstar{i - 1}"), orc.stage[i]
        assert rel(orc.residual(i), g(f"residual_{i}")) < TOL
        A = orc.tentative_matrix(Qstar, a * dt)  # :233-235
        assert rel((A @ X.ravel()).reshape(X.shape), g(f"a_tentative_{i}")) < TOL
        rhs = (orc.residual(i) - orc.mass(st["Q"])
               + a * dt * (o.f_impl_apply(st["Q"], Qstar) + o.pressure_gradient(st["p"], st["l"])))  # :239-247, as in step()
        assert rel(rhs, g(f"b_rhs_tentative_{i}")) < TOL
        Rp = -1.0 / (a * dt) * o.weak_divergence(g(f"in_Qtent{i}"))  # :177-179
        assert rel(Rp, g(f"b_rhs_mixed_poisson_{i}")) < TOL


@pytest.mark.parametrize("tag,mesh_fn,k", [("trace_square_k2", lambda: UnitSquareMesh(3, perturb=0.15), 2),
                                           ("trace_periodic_k1", lambda: PeriodicSquareMesh(3, L=2 * np.pi), 1)])
def test_reconstruct_trace_forms(tag, mesh_fn, k):
    """`_reconstruct_trace` (`hdg_imex.py:450-469`): a_trace is tau (2 | 1) |F| times the identity in the orthonormal
    Legendre basis, and the oracle's reconstructed trace solves a_trace(lambda) == b_rhs_trace"""
    mesh = mesh_fn()
    o = HDGOracle(mesh, k)
    g = lambda name: GOLD[f"{tag}/{name}"]
    mult = np.where(mesh.facet_cell[:, 1] >= 0, 2.0, 1.0)
    scale = (o.tau * mult * mesh.facet_length())[:, None]
    assert rel(scale * g("in_lam"), g("a_trace_action")) < TOL
    lam = o.reconstruct_trace(g("in_Q"), g("in_p"))
    assert rel(scale * lam, g("b_rhs_trace")) < TOL
